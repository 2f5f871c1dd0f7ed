"""GraphDef loss network (inception5h path, reference styler_base.py:17-31,53-57,91-94): the protobuf codec against
google.protobuf, the graph interpreter (engine, C-ABI kernels) against the oracle's torch interpretation on a
seeded synthetic graph with the inception5h topology and node names, and the Styler loop driving it.

The network itself is parity-UNPINNED (neither the .pb file nor TensorFlow exists here; see oracle/graphnet.py)."""
import numpy as np
import pytest
import torch

from helpers import smoke_cfg
from lnst import graphdef, synth
from lnst.graphnet import GraphNet
from oracle import graphnet as OG
from oracle import vgg as OV
from test_kernel_parity import close


# ---- wire format vs the protobuf library ---------------------------------------------------------------
def _tf_messages():
    """GraphDef / NodeDef / AttrValue / TensorProto / TensorShapeProto declared on the fly with TF's field numbers."""
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    F = descriptor_pb2.FieldDescriptorProto
    fd = descriptor_pb2.FileDescriptorProto(name='lnst_tf_subset.proto', package='lnsttf', syntax='proto3')

    def msg(name, fields, nested=()):
        m = fd.message_type.add(name=name)
        for fname, num, typ, label, tname in fields:
            f = m.field.add(name=fname, number=num, type=typ, label=label)
            if tname:
                f.type_name = '.lnsttf.' + tname
        return m

    O, R = F.LABEL_OPTIONAL, F.LABEL_REPEATED
    msg('Dim', [('size', 1, F.TYPE_INT64, O, None)])
    msg('TensorShapeProto', [('dim', 2, F.TYPE_MESSAGE, R, 'Dim')])
    msg('TensorProto', [('dtype', 1, F.TYPE_INT32, O, None), ('tensor_shape', 2, F.TYPE_MESSAGE, O, 'TensorShapeProto'),
                        ('tensor_content', 4, F.TYPE_BYTES, O, None), ('float_val', 5, F.TYPE_FLOAT, R, None),
                        ('int_val', 7, F.TYPE_INT32, R, None)])
    msg('ListValue', [('s', 2, F.TYPE_BYTES, R, None), ('i', 3, F.TYPE_INT64, R, None), ('f', 4, F.TYPE_FLOAT, R, None)])
    msg('AttrValue', [('list', 1, F.TYPE_MESSAGE, O, 'ListValue'), ('s', 2, F.TYPE_BYTES, O, None),
                      ('i', 3, F.TYPE_INT64, O, None), ('f', 4, F.TYPE_FLOAT, O, None), ('b', 5, F.TYPE_BOOL, O, None),
                      ('tensor', 8, F.TYPE_MESSAGE, O, 'TensorProto')])
    msg('AttrEntry', [('key', 1, F.TYPE_STRING, O, None), ('value', 2, F.TYPE_MESSAGE, O, 'AttrValue')])
    msg('NodeDef', [('name', 1, F.TYPE_STRING, O, None), ('op', 2, F.TYPE_STRING, O, None),
                    ('input', 3, F.TYPE_STRING, R, None), ('attr', 5, F.TYPE_MESSAGE, R, 'AttrEntry')])
    msg('GraphDef', [('node', 1, F.TYPE_MESSAGE, R, 'NodeDef')])
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    get = getattr(message_factory, 'GetMessageClass', None)
    if get is None:
        fac = message_factory.MessageFactory(pool)
        get = fac.GetPrototype
    return {n: get(pool.FindMessageTypeByName('lnsttf.' + n)) for n in ('GraphDef', 'NodeDef', 'AttrValue')}


def test_graphdef_codec_matches_protobuf_library():
    M = _tf_messages()
    # 1. what the protobuf library writes, the reader parses
    g = M['GraphDef']()
    n = g.node.add(name='conv2d0_w', op='Const')
    e = n.attr.add(key='value')
    w = np.arange(24, dtype=np.float32).reshape(2, 3, 4) - 7.5
    e.value.tensor.dtype = 1
    for d in w.shape:
        e.value.tensor.tensor_shape.dim.add(size=d)
    e.value.tensor.tensor_content = w.tobytes()
    n = g.node.add(name='axis', op='Const')
    e = n.attr.add(key='value')
    e.value.tensor.dtype = 3
    e.value.tensor.int_val.append(3)                              # scalar through int_val, empty shape
    n = g.node.add(name='fill', op='Const')
    e = n.attr.add(key='value')
    e.value.tensor.dtype = 1
    e.value.tensor.tensor_shape.dim.add(size=5)
    e.value.tensor.float_val.append(0.25)                         # one value standing for a filled tensor
    n = g.node.add(name='c', op='Conv2D')
    n.input.extend(['input', 'conv2d0_w', '^ctrl'])
    n.attr.add(key='strides').value.list.i.extend([1, 2, 2, 1])
    n.attr.add(key='padding').value.s = b'SAME'
    n.attr.add(key='alpha').value.f = 2e-5
    n.attr.add(key='depth_radius').value.i = -2
    n.attr.add(key='flag').value.b = True
    nodes = graphdef.parse(g.SerializeToString())
    assert [x.name for x in nodes] == ['conv2d0_w', 'axis', 'fill', 'c']
    np.testing.assert_array_equal(nodes[0].attr['value'], w)
    assert int(nodes[1].attr['value']) == 3 and nodes[1].attr['value'].shape == ()
    np.testing.assert_array_equal(nodes[2].attr['value'], np.full(5, 0.25, np.float32))
    c = nodes[3]
    assert c.op == 'Conv2D' and c.inputs == ['input', 'conv2d0_w', '^ctrl']
    assert c.attr['strides'] == [1, 2, 2, 1] and c.attr['padding'] == b'SAME' and c.attr['flag'] is True
    assert c.attr['depth_radius'] == -2 and abs(c.attr['alpha'] - 2e-5) < 1e-12
    # 2. what the writer emits, the protobuf library parses to the same content
    blob = graphdef.serialize(nodes)
    g2 = M['GraphDef']()
    g2.ParseFromString(blob)
    assert [x.name for x in g2.node] == ['conv2d0_w', 'axis', 'fill', 'c']
    attrs = {a.key: a.value for a in g2.node[3].attr}
    assert list(attrs['strides'].list.i) == [1, 2, 2, 1] and attrs['padding'].s == b'SAME'
    assert attrs['depth_radius'].i == -2 and abs(attrs['alpha'].f - 2e-5) < 1e-12 and attrs['flag'].b is True
    t = {a.key: a.value for a in g2.node[0].attr}['value'].tensor
    np.testing.assert_array_equal(np.frombuffer(t.tensor_content, np.float32).reshape(2, 3, 4), w)
    # 3. round trip of a whole synthetic inception graph
    big = synth.inception5h_nodes(width_div=16, upto='mixed3a')
    back = graphdef.parse(graphdef.serialize(big))
    assert [(x.name, x.op, x.inputs) for x in back] == [(x.name, x.op, x.inputs) for x in big]
    for a, b in zip(big, back):
        for k, v in a.attr.items():
            if isinstance(v, np.ndarray):
                np.testing.assert_array_equal(v, b.attr[k])
            elif isinstance(v, float):
                assert abs(v - b.attr[k]) < 1e-9
            else:
                assert v == b.attr[k], (a.name, k)


def test_synthetic_graph_has_the_reference_tensor_names():
    names = {n.name for n in synth.inception5h_nodes(width_div=16)}
    for l in ('conv2d2', 'mixed3b', 'mixed4b', 'mixed3b_3x3_bottleneck_pre_relu', 'mixed4b_pool_reduce_pre_relu',
              'mixed4d_3x3_bottleneck_pre_relu', 'conv2d0_pre_relu/conv'):     # test_smokegun.py:141, run.bat:15-20,
        assert l in names                                                       # config.py:88, styler_base.py:29
    full = {n.name: n for n in synth.inception5h_nodes(width_div=1, upto='mixed4d')}
    assert full['mixed4d_3x3_bottleneck_w'].attr['value'].shape == (1, 1, 512, 144)      # channel 139 exists (config.py:89)
    assert full['mixed3b_3x3_bottleneck_w'].attr['value'].shape[-1] == 128               # channels 44, 65 (run.bat:15-16)


# ---- interpreter vs the oracle ---------------------------------------------------------------------------
@pytest.mark.parametrize('pool1', [False, True])
def test_graphnet_forward_backward(dev, pool1):
    nodes = synth.inception5h_nodes(width_div=8, upto='mixed4a')
    hw = (40, 36) if not pool1 else (22, 20)
    img = torch.tensor(np.random.RandomState(0).uniform(0, 255, (2,) + hw + (3,)).astype(np.float32), requires_grad=True)
    # a post-ReLU tensor, a pre-ReLU one (keeps the separate Relu pass), a branch output (its module then concatenates
    # by copy), concats whose branches write in place
    layers = ['conv2d2', 'mixed3a_3x3_bottleneck_pre_relu', 'mixed3a_5x5', 'mixed3b', 'mixed4a_pool_reduce_pre_relu', 'mixed4a']
    want = OG.forward(img, nodes, layers, pool1=pool1)
    net = GraphNet(nodes, dev, pool1=pool1)
    x = OV.preprocess(img.detach()).contiguous().to(dev)
    acts = net.forward(x, layers)
    relu_of, slot = acts['__fusion__']
    assert 'mixed3b_1x1' in slot and 'mixed3a_1x1' not in slot and 'mixed3a_3x3_bottleneck' not in relu_of
    assert 'mixed3b_1x1' not in acts and 'mixed3b_1x1_pre_relu' not in acts          # never materialised on their own
    for l in layers + ['localresponsenorm0', 'maxpool1', 'mixed3a_pool', 'mixed3a']:
        close(acts[l], want[l], tol=2e-5, what=l)
    # cotangents on three tensors at different depths, one of them pre-ReLU, one a concat
    rng = np.random.RandomState(1)
    cot = {l: torch.tensor(rng.randn(*want[l].shape).astype(np.float32)) for l in
           ('conv2d2', 'mixed3a_3x3_bottleneck_pre_relu', 'mixed3a_5x5', 'mixed3b', 'mixed4a')}
    sum((want[l] * c).sum() for l, c in cot.items()).backward()

    def add(name, g):
        c = cot[name].to(dev)
        return c.clone() if g is None else g.add_(c)

    g_x = net.backward(x, acts, list(cot), add, set(cot))
    close(g_x, img.grad, tol=5e-5, what='d loss / d input')


def test_graphnet_errors(dev):
    nodes = synth.inception5h_nodes(width_div=16, upto='mixed3a')
    net = GraphNet(nodes, dev)
    x = torch.zeros(1, 16, 16, 3).to(dev)
    with pytest.raises(KeyError):
        net.forward(x, ['mixed9z'])
    from lnst.graphdef import Node
    bad = nodes + [Node('sm', 'Softmax', ['mixed3a'], {})]
    with pytest.raises(NotImplementedError):
        GraphNet(bad, dev).forward(x, ['sm'])
    assert 'mixed3a' in net.forward(x, ['import/mixed3a:0'])          # the reference's spelling (styler_base.py:94)


# ---- the Styler loop on the inception network (run.bat:15-20 semantic transfer; style layers of test_smokegun.py:141)
def _inception_cfg(**kw):
    base = dict(res=20, iter=3, network='tensorflow_inception_graph.pb', rotate=True, n_views=3,
                w_style=0, w_content=1, content_layer='mixed3b_3x3_bottleneck_pre_relu', content_channel=5)
    base.update(kw)
    return smoke_cfg(**base)


@pytest.mark.parametrize('mode', ['semantic', 'style', 'both'])
def test_styler_with_inception_matches_oracle(dev, mode):
    from lnst.styler_3p import Styler
    from oracle.styler import Oracle3P
    kw = {}
    if mode in ('style', 'both'):
        kw.update(w_style=1, style_layer=['conv2d2', 'mixed3a', 'mixed3b'], w_style_layer=[1, 1, 1])
    if mode == 'style':
        kw.update(w_content=0)
    nodes = synth.inception5h_nodes(width_div=8, upto='mixed3b')
    p, r = synth.smoke_particles(900, 2, pad=3)
    sty = synth.style_image(20, 20)
    new = Styler(_inception_cfg(**kw), weights=nodes, device=dev)
    new.style_img = sty if mode != 'semantic' else None
    out = new.run({'p': p, 'r': r})
    ref = Oracle3P(_inception_cfg(**kw), nodes).run({'p': p, 'r': r}, style_targets=[sty] if mode != 'semantic' else None)
    np.testing.assert_allclose(out['l'][0], ref['l'][0], rtol=3e-4)
    err = np.abs(out['d'] - ref['d']).max() / np.abs(ref['d']).max()
    assert err < 3e-4, err


def test_styler_reads_the_pb_file(dev, tmp_path):
    """config.network / data_dir / model_dir -> <data_dir>/<model_dir>/tensorflow_inception_graph.pb (styler_base.py:18-23)"""
    from lnst.styler_3p import Styler
    nodes = synth.inception5h_nodes(width_div=16, upto='mixed3a')
    cfg = _inception_cfg(iter=1, rotate=False, content_layer='mixed3a_1x1_pre_relu', content_channel=0)
    cfg.data_dir, cfg.model_dir = str(tmp_path), 'model'
    (tmp_path / 'model').mkdir()
    (tmp_path / 'model' / cfg.network).write_bytes(graphdef.serialize(nodes))
    p, r = synth.smoke_particles(300, 2)
    a = Styler(cfg, device=dev).run({'p': p, 'r': r})
    b = Styler(cfg, weights=nodes, device=dev).run({'p': p, 'r': r})
    np.testing.assert_allclose(a['l'][0], b['l'][0], rtol=1e-4)


# ---- the reference's own inception path, run here (tests/golden/make_reference_golden.py: the unmodified
# styler_base / styler_3p parse the GraphDef file, import it and read layers by tensor name on oracle/tfshim) --------
REF_CASES = ['density_inception', 'density_inception_pool1', 'density_inception_logits']


def _ref_setup(name):
    import os
    import sys
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
    sys.path.insert(0, gold)
    import make_reference_golden as M
    import test_reference_golden as TR
    cfg, params = M.case_inputs(name)
    return M, TR, cfg, params, dict(np.load(os.path.join(gold, 'ref_%s.npz' % name)))


@pytest.mark.parametrize('name', REF_CASES)
def test_oracle_matches_reference_inception_run(name):
    from oracle.styler import Oracle3P
    M, TR, cfg, params, want = _ref_setup(name)
    out = Oracle3P(cfg, M.inception_nodes(head='logits' in name)).run(
        params, style_targets=TR._style_targets(cfg), content_targets=TR._content_targets(cfg), view_mode='sequential')
    TR._check(out, want, '3d', ltol=2e-5, ftol=1e-4)


@pytest.mark.parametrize('name', REF_CASES)
def test_engine_matches_reference_inception_run(name, dev):
    from lnst.styler_3p import Styler
    M, TR, cfg, params, want = _ref_setup(name)
    cfg.view_mode = 'sequential'
    st = Styler(cfg, weights=M.inception_nodes(head='logits' in name), device=dev)
    tg, ct = TR._style_targets(cfg), TR._content_targets(cfg)
    if tg is not None:
        st.style_img = tg[0]
    if ct is not None:
        st.content_img = ct[0]
    TR._check(st.run(params), want, '3d')



def test_graphnet_classifier_head(dev):
    """avgpool0 -> reshape [-1, C] -> MatMul -> BiasAdd (softmax2_pre_activation): forward and data gradient vs the oracle"""
    nodes = synth.inception5h_nodes(width_div=8, upto='mixed3b', head_pool=2)
    img = torch.tensor(np.random.RandomState(4).uniform(0, 255, (2, 36, 30, 3)).astype(np.float32), requires_grad=True)
    want = OG.forward(img, nodes, ['softmax2_pre_activation'])
    net = GraphNet(nodes, dev)
    x = OV.preprocess(img.detach()).contiguous().to(dev)
    acts = net.forward(x, ['softmax2_pre_activation'])
    logits = acts['softmax2_pre_activation']                       # [n, rows per image, 1, classes]
    assert logits.shape[0] == 2 and logits.shape[2] == 1
    close(logits.reshape(-1, logits.shape[-1]), want['softmax2_pre_activation'], tol=2e-5, what='logits')
    close(acts['avgpool0'], want['avgpool0'], tol=2e-5, what='avgpool0')
    cot = torch.tensor(np.random.RandomState(5).randn(*want['softmax2_pre_activation'].shape).astype(np.float32))
    (want['softmax2_pre_activation'] * cot).sum().backward()
    g = cot.reshape(logits.shape).to(dev)
    g_x = net.backward(x, acts, ['softmax2_pre_activation'], lambda n, gg: g.clone(), {'softmax2_pre_activation'})
    close(g_x, img.grad, tol=5e-5, what='d logits / d input')
