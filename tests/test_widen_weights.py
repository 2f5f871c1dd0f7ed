"""Loss-network weight loaders (``lnst/vgg.py``): slim ``.npz`` export and torchvision state dicts."""
import os

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


# ---- TensorFlow V1 checkpoint files (lnst/tfckpt.py) ------------------------------------------------------
def _slim_tensors(model='vgg_16', with_extra=True):
    w = OV.synthetic_weights(model)
    t = {}
    for name, (wt, b) in w.items():
        t['%s/%s/%s/weights' % (model, name.split('_')[0], name)] = wt.numpy()
        t['%s/%s/%s/biases' % (model, name.split('_')[0], name)] = b.numpy()
    if with_extra:                                             # the zoo file also holds the classifier and a step
        t['%s/fc8/biases' % model] = np.arange(7, dtype=np.float32)
        t['global_step'] = np.asarray(123, np.int32)
    return w, t


def test_v1_checkpoint_round_trip_and_loader(tmp_path):
    from lnst import tfckpt
    w, t = _slim_tensors()
    path = str(tmp_path / 'vgg_16.ckpt')
    tfckpt.write(path, t)
    back = tfckpt.read(path)
    assert sorted(back) == sorted(t)
    for k in t:
        np.testing.assert_array_equal(back[k], t[k])
        assert back[k].shape == np.asarray(t[k]).shape
    got = vgg.load_weights(path, 'vgg_16')                      # config.network points at the .ckpt itself
    for name in w:
        torch.testing.assert_close(got[name][0], w[name][0], rtol=0, atol=0)
        torch.testing.assert_close(got[name][1], w[name][1], rtol=0, atol=0)
    only = tfckpt.read(path, names={'global_step'})
    assert list(only) == ['global_step'] and int(only['global_step']) == 123


def test_v1_checkpoint_table_layout(tmp_path):
    """footer magic, block trailers with valid masked crc32c, prefix-compressed keys and snappy blocks are read"""
    import struct
    from lnst import tfckpt
    assert tfckpt.crc32c(b'123456789') == 0xE3069283            # the CRC-32C check value
    _, t = _slim_tensors(with_extra=False)
    path = str(tmp_path / 'a.ckpt')
    tfckpt.write(path, {k: t[k] for k in list(t)[:4]})
    blob = open(path, 'rb').read()
    assert struct.unpack('<Q', blob[-8:])[0] == 0xdb4775248b80fb57
    foot = memoryview(blob[-48:])
    mo, ms, pos = tfckpt._handle(foot, 0)
    io, isz, _ = tfckpt._handle(foot, pos)
    for off, size in ((mo, ms), (io, isz)):
        assert blob[off + size] == 0
        crc, = struct.unpack_from('<I', blob, off + size + 1)
        assert crc == tfckpt._mask(tfckpt.crc32c(b'\x00', tfckpt.crc32c(blob[off:off + size])))
    # a block with shared key prefixes and restart points, as TF's TableBuilder writes them
    ents = [(b'abc', b'1'), (b'abd', b'22'), (b'abdz', b''), (b'b', b'4444')]
    body = bytearray()
    prev = b''
    for i, (k, v) in enumerate(ents):
        shared = 0 if i % 2 == 0 else len(os.path.commonprefix([prev, k]))
        body += bytes([shared, len(k) - shared, len(v)]) + k[shared:] + v
        prev = k
    # restart offsets of entries 0 and 2
    o2 = len(bytes([0, 3, 1]) + b'abc' + b'1') + len(bytes([2, 1, 2]) + b'd' + b'22')
    body += struct.pack('<III', 0, o2, 2)
    raw = bytes(body) + b'\x00'
    assert [(k, bytes(v)) for k, v in tfckpt._block(memoryview(raw), 0, len(body))] == ents
    # the same block snappy-compressed by hand: one literal, then a back-reference copy
    data = b'abcdabcdabcdXY'
    comp = bytes([len(data)]) + bytes([(4 - 1) << 2]) + b'abcd' + bytes([((8 - 4) << 2) | 1 | (0 << 5), 4]) + \
        bytes([(2 - 1) << 2]) + b'XY'
    assert bytes(tfckpt._snappy(memoryview(comp))) == data


def test_v1_checkpoint_partitioned_tensor_is_reassembled(tmp_path):
    """two slices of one tensor (a partitioned variable) land in the right rows"""
    from lnst import tfckpt
    from lnst.graphdef import _enc_varint, _ld, _enc_shape
    full = np.arange(12, dtype=np.float32).reshape(4, 3)

    def slice_entry(rows):
        a = full[rows[0]:rows[0] + rows[1]]
        ext = _ld(1, _enc_varint(1 << 3) + _enc_varint(rows[0]) + _enc_varint(2 << 3) + _enc_varint(rows[1])) + _ld(1, b'')
        tp = _enc_varint(1 << 3) + _enc_varint(1) + _ld(2, _enc_shape(a.shape)) + _ld(5, a.tobytes())
        return _ld(2, _ld(1, b'v') + _ld(2, ext) + _ld(3, tp))

    meta = _ld(1, _ld(1, _ld(1, b'v') + _ld(2, _enc_shape(full.shape)) + _enc_varint(3 << 3) + _enc_varint(1)))
    entries = [(b'', meta), (b'\x00v\x00\x01a', slice_entry((0, 1))), (b'\x00v\x00\x01b', slice_entry((1, 3)))]
    out, index = bytearray(), []
    for k, v in entries:
        blk = tfckpt._enc_block([(k, v)])
        index.append((k + b'\x00', _enc_varint(len(out)) + _enc_varint(len(blk))))
        out += blk + b'\x00' + b'\x00' * 4
    mo = len(out)
    mblk = tfckpt._enc_block([])
    out += mblk + b'\x00' * 5
    io = len(out)
    iblk = tfckpt._enc_block(index)
    out += iblk + b'\x00' * 5
    import struct
    foot = _enc_varint(mo) + _enc_varint(len(mblk)) + _enc_varint(io) + _enc_varint(len(iblk))
    out += foot + b'\x00' * (40 - len(foot)) + struct.pack('<Q', tfckpt.MAGIC)
    p = tmp_path / 'part.ckpt'
    p.write_bytes(bytes(out))
    np.testing.assert_array_equal(tfckpt.read(str(p))['v'], full)
    with pytest.raises(ValueError):
        bad = tmp_path / 'bad.ckpt'
        bad.write_bytes(b'x' * 100)
        tfckpt.read(str(bad))


def _slice_messages():
    """SavedTensorSlices & co. declared on the fly with TensorFlow's field numbers (saved_tensor_slice.proto,
    tensor_slice.proto, tensor.proto, tensor_shape.proto) -- the protobuf library as an independent codec."""
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    Fd = descriptor_pb2.FieldDescriptorProto
    fd = descriptor_pb2.FileDescriptorProto(name='lnst_ckpt_subset.proto', package='lnstck', syntax='proto3')
    O, R = Fd.LABEL_OPTIONAL, Fd.LABEL_REPEATED

    def msg(name, spec):
        m = fd.message_type.add(name=name)
        for fname, num, typ, label, tname in spec:
            f = m.field.add(name=fname, number=num, type=typ, label=label)
            if tname:
                f.type_name = '.lnstck.' + tname

    msg('Dim', [('size', 1, Fd.TYPE_INT64, O, None)])
    msg('TensorShapeProto', [('dim', 2, Fd.TYPE_MESSAGE, R, 'Dim')])
    msg('TensorProto', [('dtype', 1, Fd.TYPE_INT32, O, None), ('tensor_shape', 2, Fd.TYPE_MESSAGE, O, 'TensorShapeProto'),
                        ('tensor_content', 4, Fd.TYPE_BYTES, O, None), ('float_val', 5, Fd.TYPE_FLOAT, R, None),
                        ('int_val', 7, Fd.TYPE_INT32, R, None)])
    msg('Extent', [('start', 1, Fd.TYPE_INT64, O, None), ('length', 2, Fd.TYPE_INT64, O, None)])
    msg('TensorSliceProto', [('extent', 1, Fd.TYPE_MESSAGE, R, 'Extent')])
    msg('SavedSliceMeta', [('name', 1, Fd.TYPE_STRING, O, None), ('shape', 2, Fd.TYPE_MESSAGE, O, 'TensorShapeProto'),
                           ('type', 3, Fd.TYPE_INT32, O, None), ('slice', 4, Fd.TYPE_MESSAGE, R, 'TensorSliceProto')])
    msg('SavedTensorSliceMeta', [('tensor', 1, Fd.TYPE_MESSAGE, R, 'SavedSliceMeta')])
    msg('SavedSlice', [('name', 1, Fd.TYPE_STRING, O, None), ('slice', 2, Fd.TYPE_MESSAGE, O, 'TensorSliceProto'),
                       ('data', 3, Fd.TYPE_MESSAGE, O, 'TensorProto')])
    msg('SavedTensorSlices', [('meta', 1, Fd.TYPE_MESSAGE, O, 'SavedTensorSliceMeta'), ('data', 2, Fd.TYPE_MESSAGE, O, 'SavedSlice')])
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    get = getattr(message_factory, 'GetMessageClass', None)
    if get is None:
        get = message_factory.MessageFactory(pool).GetPrototype
    return get(pool.FindMessageTypeByName('lnstck.SavedTensorSlices'))


def test_v1_checkpoint_protos_against_the_protobuf_library(tmp_path):
    import struct
    from lnst import tfckpt
    from lnst.graphdef import _enc_varint
    STS = _slice_messages()
    # 1. what lnst.tfckpt.write emits parses with the protobuf library to the same content
    t = {'a/weights': np.arange(24, dtype=np.float32).reshape(2, 3, 4) - 3.5, 'step': np.asarray([7, -2], np.int32)}
    path = str(tmp_path / 'w.ckpt')
    tfckpt.write(path, t)
    seen = {}
    blob = memoryview(open(path, 'rb').read())
    for key, val in tfckpt._entries(blob):
        m = STS()
        m.ParseFromString(bytes(val))
        if key == b'':
            meta = {s.name: ([d.size for d in s.shape.dim], s.type, len(s.slice)) for s in m.meta.tensor}
        else:
            d = m.data
            seen[d.name] = (list(d.data.float_val) or list(d.data.int_val), [d_.size for d_ in d.data.tensor_shape.dim],
                            [(e.start, e.length) for e in d.slice.extent])
    assert meta == {'a/weights': ([2, 3, 4], 1, 1), 'step': ([2], 3, 1)}
    assert seen['a/weights'][0] == list(t['a/weights'].reshape(-1)) and seen['a/weights'][2] == [(0, 0)] * 3
    assert seen['step'][0] == [7, -2]
    # 2. protos written by the protobuf library (packed float_val, one full slice and one two-part tensor) are read back
    m0 = STS()
    for name, shape, dt in (('x', [2, 2], 1), ('y', [4], 1)):
        s = m0.meta.tensor.add(name=name, type=dt)
        for d in shape:
            s.shape.dim.add(size=d)
    ents = [(b'', m0.SerializeToString())]
    mx = STS()
    mx.data.name = 'x'
    mx.data.slice.extent.add()
    mx.data.slice.extent.add()
    mx.data.data.dtype = 1
    mx.data.data.float_val.extend([1.5, 2.5, 3.5, 4.5])
    ents.append((b'\x00x\x00\x01', mx.SerializeToString()))
    for start, vals in ((0, [10.0]), (1, [11.0, 12.0, 13.0])):
        my = STS()
        my.data.name = 'y'
        e = my.data.slice.extent.add(start=start, length=len(vals))
        my.data.data.dtype = 1
        my.data.data.float_val.extend(vals)
        ents.append((b'\x00y\x00\x01' + bytes([start]), my.SerializeToString()))
    out, index = bytearray(), []
    blk = tfckpt._enc_block(ents)                               # all entries in ONE data block this time
    index.append((ents[-1][0] + b'\x00', _enc_varint(0) + _enc_varint(len(blk))))
    out += blk + b'\x00' * 5
    mo, mblk = len(out), tfckpt._enc_block([])
    out += mblk + b'\x00' * 5
    io, iblk = len(out), tfckpt._enc_block(index)
    out += iblk + b'\x00' * 5
    foot = _enc_varint(mo) + _enc_varint(len(mblk)) + _enc_varint(io) + _enc_varint(len(iblk))
    out += foot + b'\x00' * (40 - len(foot)) + struct.pack('<Q', tfckpt.MAGIC)
    p2 = tmp_path / 'pb.ckpt'
    p2.write_bytes(bytes(out))
    got = tfckpt.read(str(p2))
    np.testing.assert_array_equal(got['x'], np.array([[1.5, 2.5], [3.5, 4.5]], np.float32))
    np.testing.assert_array_equal(got['y'], np.array([10, 11, 12, 13], np.float32))
