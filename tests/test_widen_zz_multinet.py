"""Multi-net loss (BASELINE.json configs[4]: inception semantic + VGG-19 style; an engine extension -- the reference
builds one network per run): content term on a GraphDef network, style term on VGG (fp32 and, on the B200, the
tcgen05 path)."""
import numpy as np
import pytest

from helpers import smoke_cfg
from lnst import synth


@pytest.mark.parametrize('math', ['fp32', pytest.param('bf16', marks=pytest.mark.gpu)])
def test_multi_net_loss_matches_oracle(dev, math):
    from lnst.styler_3p import Styler
    from oracle.styler import Oracle3P
    import oracle.vgg
    if math == 'bf16' and dev.type != 'cuda':
        pytest.skip('tensor-core path needs the GPU')
    kw = dict(res=20, iter=3, rotate=True, n_views=3, network='vgg_19.ckpt', conv_math=math,
              style_layer=['conv1_2', 'conv2_1'], w_style_layer=[0.5, 0.5], w_style=1.0,
              content_network='tensorflow_inception_graph.pb', w_content=50.0,
              content_layer='mixed3a_3x3_bottleneck_pre_relu', content_channel=3)
    nodes = synth.inception5h_nodes(width_div=8, upto='mixed3a')
    p, r = synth.smoke_particles(900, 2, pad=3)
    sty = synth.style_image(20, 20)
    new = Styler(smoke_cfg(**kw), weights=synth.vgg_weights(), content_weights=nodes, device=dev)
    new.style_img = sty
    out = new.run({'p': p, 'r': r})
    ref = Oracle3P(smoke_cfg(**kw), oracle.vgg.synthetic_weights(), content_weights=nodes).run(
        {'p': p, 'r': r}, style_targets=[sty])
    tol = 3e-4 if math == 'fp32' else 5e-2
    np.testing.assert_allclose(out['l'][0], ref['l'][0], rtol=tol)
    err = np.abs(out['d'] - ref['d']).max() / np.abs(ref['d']).max()
    assert err < tol, err
    # both terms are live: dropping either one changes the loss
    only_style = Oracle3P(smoke_cfg(**dict(kw, w_content=0)), oracle.vgg.synthetic_weights()).run(
        {'p': p, 'r': r}, style_targets=[sty])
    assert abs(only_style['l'][0][0] - ref['l'][0][0]) > 1e-3 * abs(ref['l'][0][0])
