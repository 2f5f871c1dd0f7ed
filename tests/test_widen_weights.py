"""Loss-network weight loaders (``lnst/vgg.py``): slim ``.npz`` export and torchvision state dicts."""
import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from lnst import vgg
from oracle import vgg as OV

_TV_CFG = {'vgg_19': [64, 64, 'M', 128, 128, 'M', 256, 256, 256, 256, 'M', 512, 512, 512, 512, 'M', 512, 512, 512, 512, 'M'],
           'vgg_16': [64, 64, 'M', 128, 128, 'M', 256, 256, 256, 'M', 512, 512, 512, 'M', 512, 512, 512, 'M']}


def _torchvision_features(model, seed=0, width_div=8):
    """torchvision.models.vgg.make_layers' module layout (conv, relu, ..., pool indices), narrow channels."""
    torch.manual_seed(seed)
    layers, cin = [], 3
    for v in _TV_CFG[model]:
        if v == 'M':
            layers.append(nn.MaxPool2d(2, 2))
        else:
            layers += [nn.Conv2d(cin, v // width_div, 3, padding=1), nn.ReLU(inplace=False)]
            cin = v // width_div
    return nn.Sequential(*layers)


@pytest.mark.parametrize('model', ['vgg_19', 'vgg_16'])
def test_torchvision_state_dict_remap(tmp_path, model):
    feats = _torchvision_features(model)
    sd = {'features.' + k: v for k, v in feats.state_dict().items()}
    sd['classifier.0.weight'] = torch.zeros(4, 4)             # ignored
    path = tmp_path / (model + '.pth')
    torch.save(sd, str(path))
    w = vgg.load_weights(str(tmp_path / (model + '.ckpt')), model)       # the reference's path (config.network)
    names = [n for n in vgg.layer_order(model) if n.startswith('conv')]
    assert list(w) == names and w['conv1_1'][0].shape[:3] == (3, 3, 3)
    # torchvision's own pipeline (x/255, mean/std normalisation) with the reference's average pooling ...
    img = torch.tensor(np.random.RandomState(1).uniform(0, 255, (1, 32, 32, 3)).astype(np.float32))
    mean, std = torch.tensor([0.485, 0.456, 0.406]), torch.tensor([0.229, 0.224, 0.225])
    x = ((img / 255 - mean) / std).permute(0, 3, 1, 2)
    want = {}
    it = iter(names)
    for m in feats:
        if isinstance(m, nn.MaxPool2d):
            if x.shape[-1] < 2:
                break                                          # the last pool (after conv5_x) is never an end point here
            x = F.avg_pool2d(x, 2, 2)
        else:
            x = m(x)
            if isinstance(m, nn.ReLU):
                want[next(it)] = x.permute(0, 2, 3, 1)
    # ... equals the slim-convention network (x - 255*mean, no std: vgg.py:50-53) on the remapped weights
    got = OV.forward(img, w, model)
    for n in ('conv1_1', 'conv2_1', 'conv3_1', names[-1]):
        torch.testing.assert_close(got[n], want[n].detach(), rtol=2e-4, atol=2e-5)


def test_npz_export_both_key_styles(tmp_path):
    w = OV.synthetic_weights('vgg_16')
    blob = {}
    for i, (name, (wt, b)) in enumerate(w.items()):
        key = 'vgg_16/%s/%s' % (name.split('_')[0], name) if i % 2 else name
        blob[key + '/weights'], blob[key + '/biases'] = wt.numpy(), b.numpy()
    np.savez(str(tmp_path / 'vgg_16.npz'), **blob)
    got = vgg.load_weights(str(tmp_path / 'vgg_16.ckpt'), 'vgg_16')
    for name in w:
        torch.testing.assert_close(got[name][0], w[name][0], rtol=0, atol=0)


def test_missing_weights_message(tmp_path):
    with pytest.raises(FileNotFoundError, match='npz'):
        vgg.load_weights(str(tmp_path / 'vgg_19.ckpt'))
    feats = _torchvision_features('vgg_16')
    with pytest.raises(ValueError, match='conv layers'):
        vgg.from_torchvision({'features.' + k: v for k, v in feats.state_dict().items()}, 'vgg_19')
