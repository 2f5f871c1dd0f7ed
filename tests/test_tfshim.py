"""The TensorFlow-1.15 stand-in (``oracle/tfshim``) checked on its own: every rule its header says it restates
explicitly, against closed forms, torch, or the independently written oracle functions.  It is the tool that
produced ``tests/golden/ref_*.npz``; these tests need neither the reference nor a GPU."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, 'oracle', 'tfshim')


@pytest.fixture(scope='module')
def tf():
    sys.path.insert(0, SHIM)
    try:
        import tensorflow as tf
        assert 'shim' in tf.__version__
        yield tf
    finally:
        sys.path.remove(SHIM)
        sys.modules.pop('tensorflow', None)


def grad_of(tf, sess, build, x0):
    """d sum(build(x)) / dx at x0 through tf.gradients."""
    ph = tf.placeholder(tf.float32, list(np.shape(x0)))
    y = build(ph)
    leaf = torch.tensor(np.asarray(x0, np.float32), requires_grad=True)
    return sess.run([y, tf.gradients(tf.reduce_sum(y), [ph])[0]], {ph: leaf})


def test_maximum_minimum_clip_gradients_follow_tf(tf):
    sess = tf.Session()
    x = np.array([-2.0, -1.0, -0.5, 0.0, 1.0, 1.5, np.nan], np.float32)
    y, g = grad_of(tf, sess, lambda t: tf.clip_by_value(t, -1, 1), x)
    np.testing.assert_array_equal(y[:-1], [-1, -1, -0.5, 0, 1, 1])
    assert y[-1] == 1.0                                    # maximum(minimum(NaN, 1), -1): the non-NaN operand wins
    np.testing.assert_array_equal(g, [0, 1, 1, 1, 1, 0, 0])    # inclusive at both bounds, zero for NaN
    y, g = grad_of(tf, sess, lambda t: tf.maximum(t, 0), np.array([-1.0, 0.0, 2.0], np.float32))
    np.testing.assert_array_equal(g, [0, 1, 1])            # passes at equality (x >= y)


def test_reduce_max_gradient_is_split_among_ties(tf):
    sess = tf.Session()
    y, g = grad_of(tf, sess, lambda t: tf.reduce_max(t), np.array([[1.0, 3.0], [3.0, 2.0]], np.float32))
    assert y == 3.0
    np.testing.assert_array_equal(g, [[0, 0.5], [0.5, 0]])


def test_scatter_nd_sums_duplicates_and_drops_out_of_range(tf):
    sess = tf.Session()
    idx = np.array([[0, 1], [0, 1], [1, 2], [2, 0], [0, -1]], np.int32)
    out = sess.run(tf.scatter_nd(idx, np.array([1.0, 2.0, 3.0, 4.0, 5.0], np.float32), [2, 3]))
    np.testing.assert_array_equal(out, [[0, 3, 0], [0, 0, 3]])


def test_linspace_and_cumsum_and_where(tf):
    sess = tf.Session()
    ls = sess.run(tf.linspace(-1.0, 1.0, 7))
    step = np.float32(2.0) / np.float32(6.0)
    np.testing.assert_array_equal(ls, np.float32(-1.0) + step * np.arange(7, dtype=np.float32))
    x = np.arange(6, dtype=np.float32).reshape(1, 3, 2)
    ph = tf.placeholder(tf.float32, [1, 3, 2])
    rev = sess.run(tf.cumsum(ph[:, ::-1], axis=1)[:, ::-1], {ph: x})         # styler_3p.py:155
    np.testing.assert_array_equal(rev[0, :, 0], [6, 6, 4])
    q = np.array([0.25, 0.75, 1.5], np.float32)
    qc = tf.constant(q)
    w = sess.run(tf.compat.v1.where(qc > 1, tf.zeros_like(qc), qc * 2))
    np.testing.assert_array_equal(w, [0.5, 1.5, 0.0])


def test_legacy_resizes_match_the_oracle_restatements(tf):
    """Two independent restatements of resize_bilinear_op.cc / resize_bicubic_op.cc (legacy pixel mapping)."""
    from oracle import render as R, loss as L
    sess = tf.Session()
    x = torch.rand(2, 9, 7, 3, generator=torch.Generator().manual_seed(1))
    ph = tf.placeholder(tf.float32, [None, None, None, 3])
    for (oh, ow) in [(13, 10), (4, 3), (9, 7), (18, 21)]:
        bil = sess.run(tf.compat.v1.image.resize(ph, (oh, ow), method=tf.image.ResizeMethod.BILINEAR), {ph: x})
        np.testing.assert_allclose(bil, R.resize_bilinear_legacy(x, oh, ow).numpy(), rtol=0, atol=2e-6)
        bic = sess.run(tf.compat.v1.image.resize(ph, (oh, ow), method=tf.image.ResizeMethod.BICUBIC), {ph: x})
        np.testing.assert_allclose(bic, L.bicubic_legacy(x, oh, ow).numpy(), rtol=0, atol=2e-3)   # 1024-entry table


def test_slim_vgg_layers_and_total_variation(tf):
    sess = tf.Session()
    slim = tf.contrib.slim
    g = torch.Generator().manual_seed(2)
    x = torch.randn(1, 8, 6, 3, generator=g)
    w, b = torch.randn(3, 3, 3, 5, generator=g), torch.randn(5, generator=g)
    ph = tf.placeholder(tf.float32, [None, None, None, 3])
    with slim.arg_scope([slim.conv2d], activation_fn=tf.nn.relu, biases_initializer=tf.zeros_initializer()):
        with tf.compat.v1.variable_scope('net'):
            y = slim.conv2d(ph, 5, [3, 3], scope='c1')
            p = slim.avg_pool2d(y, [2, 2], scope='p1')
    tf.register_checkpoint('ck', {'net/c1/weights': w.numpy(), 'net/c1/biases': b.numpy()})
    slim.assign_from_checkpoint_fn('ck', slim.get_model_variables('net'))(sess)
    yv, pv = sess.run([y, p], {ph: x})
    want = torch.relu(torch.nn.functional.conv2d(x.permute(0, 3, 1, 2), w.permute(3, 2, 0, 1), b, padding=1))
    np.testing.assert_allclose(yv, want.permute(0, 2, 3, 1).numpy(), atol=1e-5)
    np.testing.assert_allclose(pv, torch.nn.functional.avg_pool2d(want, 2, 2).permute(0, 2, 3, 1).numpy(), atol=1e-5)
    assert y.shape[-1] == 5 and p.shape[-1] == 5                       # static channel count (styler_base.py:99)
    tv = sess.run(tf.compat.v1.image.total_variation(ph), {ph: x})
    want_tv = (x[:, 1:] - x[:, :-1]).abs().sum() + (x[:, :, 1:] - x[:, :, :-1]).abs().sum()
    np.testing.assert_allclose(tv, [want_tv.item()], rtol=1e-6)


def test_adam_optimizer_is_tf_apply_adam_with_feeds_and_variable_reinit(tf):
    """The session pattern of styler_3p.py:307-334: variable initialised from a placeholder every step, slots and
    beta powers persistent, NaN gradients keep m, v and the variable NaN."""
    from oracle.adam import TFAdam
    sess = tf.Session()
    ph = tf.placeholder(tf.float32, [None, 2])
    var = tf.Variable(ph, validate_shape=False)
    v_ = tf.reshape(var, tf.shape(ph))
    target = tf.placeholder(tf.float32, [None, 2])
    loss = tf.reduce_sum(tf.math.squared_difference(tf.clip_by_value(v_, -1, 1), target) * tf.sqrt(tf.abs(v_)))
    lr = tf.placeholder(tf.float32)
    opt = tf.compat.v1.train.AdamOptimizer(learning_rate=lr)
    train = opt.minimize(loss, var_list=[var])
    rng = np.random.RandomState(0)
    x = rng.uniform(-0.9, 0.9, (5, 2)).astype(np.float32)
    x[1, 1] = 0.0                                                       # sqrt'(0)*0 -> NaN gradient
    tg = rng.randn(5, 2).astype(np.float32)
    feed = {ph: x, target: tg, lr: 0.05}
    sess.run(tf.compat.v1.initializers.variables([var]), feed)
    sess.run(tf.compat.v1.variables_initializer(opt.variables()), feed)
    ref = TFAdam()
    xr = torch.tensor(x)
    for it in range(3):
        sess.run(tf.compat.v1.initializers.variables([var]), feed)    # re-assign from the host iterate
        _, l = sess.run([train, loss], feed)
        got = sess.run(var, feed)
        xt = xr.clone().requires_grad_(True)
        lt = (((torch.fmax(torch.fmin(xt, torch.tensor(1.0)), torch.tensor(-1.0)) - torch.tensor(tg)) ** 2)
              * torch.sqrt(xt.abs())).sum()
        (gr,) = torch.autograd.grad(lt, xt)
        want = ref.step(xr, gr, 0.05)
        assert np.isnan(got[1, 1]) and torch.isnan(want[1, 1])
        ok = ~np.isnan(got)
        np.testing.assert_allclose(got[ok], want.numpy()[ok], rtol=2e-5, atol=1e-6)   # Adam's first steps amplify rounding
        np.testing.assert_allclose(l, float(lt), rtol=1e-6)
        x = np.nan_to_num(got)                                          # the host copy (styler_3p.py:360)
        xr = torch.tensor(x)
        feed[ph] = x


@pytest.mark.skipif(not os.path.isdir(os.environ.get('LNST_REFERENCE', '/root/reference')),
                    reason='the reference sources exist only in the build container')
def test_committed_reference_vectors_are_reproducible():
    """Re-runs one loop-level case and the operator vectors through the reference + stand-in (separate process,
    so the stand-in never enters this interpreter's module table) and compares with the committed files."""
    import subprocess
    gold = os.path.join(ROOT, 'tests', 'golden')
    code = ("import sys, numpy as np; sys.path.insert(0, %r); import make_reference_golden as M;"
            "o = M.run_reference('density_noview'); w = np.load(%r);"
            "assert all(np.array_equal(o[k], w[k]) for k in w.files), 'loop vectors differ';"
            "o = M.run_ops(); w = np.load(%r);"
            "assert all(np.array_equal(o[k], w[k], equal_nan=True) for k in w.files), 'operator vectors differ'"
            % (gold, os.path.join(gold, 'ref_density_noview.npz'), os.path.join(gold, 'ref_ops.npz')))
    r = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
