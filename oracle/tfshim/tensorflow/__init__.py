"""A stand-in for the subset of TensorFlow 1.15 the reference's hot path calls.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``): imported only by
``tests/golden/make_reference_golden.py`` to EXECUTE THE UNMODIFIED REFERENCE SOURCES
(``/root/reference/{styler_3p,styler_2p,styler_base,transform,vgg}.py``) in this container and record
their outputs as golden vectors.  Nothing in the product or in the GPU tests imports it.

Why it exists.  The reference's only backend is TensorFlow 1.15 (+ ``tf.contrib.slim``), which
has no wheel for this image's Python 3.12 and cannot be installed offline.  TensorFlow is thus a
third-party dependency absent from ``/root/reference``; this module restates the *published
semantics* of every TF op the path touches (forward value, gradient rule, dtype conversion, static
shape where the reference reads it) on top of torch-CPU tensors, so that the reference's own graph
construction code, session loop and optimiser calls run line by line.  What is pinned by a run is
the composition -- which ops, in which order, with which arguments, on which feeds: exactly the
part a hand restatement (``oracle/``) can get wrong.  What is NOT pinned is TF's floating-point
summation order inside a kernel (conv, matmul, reduce) -- covered by the stated fp32 tolerance.

Model: a lazy dataflow graph.  Every ``tf.*`` call returns a ``Tensor`` node (closure + inputs);
``Session.run(fetches, feed_dict)`` evaluates the needed nodes once (memo per run) with torch
autograd recording, then runs fetched ``Operation``s (variable initialisers, Adam ``minimize``
steps) against the session's variable store.  Any node may be fed (``feed_dict`` overrides it), as
TF allows.

Op semantics that differ from torch's defaults and are restated explicitly (TF 1.15 sources):
  * ``maximum/minimum`` gradient goes to x where ``x >= y`` / ``x <= y`` (``math_grad.py
    _MaximumMinimumGrad``); ``clip_by_value`` = ``maximum(minimum(x, hi), lo)`` (``clip_ops.py``);
  * ``reduce_max/min`` gradient is split equally among ties (``math_grad.py _MinOrMaxGrad``);
  * ``scatter_nd`` sums duplicates; out-of-range indices are dropped (GPU kernel behaviour -- the
    CPU kernel raises; the reference is written for the GPU);
  * ``linspace`` = ``start + i*step`` in fp32 (``sequence_ops.cc`` LinSpaceOp, 1.15);
  * ``image.resize`` v1 (``align_corners=False``, no half-pixel centres): bilinear
    ``in = out*scale, lo = floor(in), hi = min(ceil(in), n-1)`` (``resize_bilinear_op.cc``);
    bicubic with A = -0.75, 1024-entry coefficient table, clamped taps (``resize_bicubic_op.cc``);
  * ``train.AdamOptimizer`` = the dense ``ApplyAdam`` CPU kernel (``training_ops.cc``):
    ``alpha = lr*sqrt(1-b2^t)/(1-b1^t); m += (g-m)(1-b1); v += (g*g-v)(1-b2);
    var -= m*alpha/(sqrt(v)+eps)``, beta powers multiplied afterwards;
  * python scalars / numpy arrays combined with a Tensor are converted to the Tensor's dtype.
``slim.assign_from_checkpoint_fn`` cannot read a TF checkpoint; it assigns from the name->array
dict registered with ``register_checkpoint(path, dict)`` (slim variable names).
"""
import builtins as _bi
import contextlib
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

sys.setrecursionlimit(20000)
__version__ = '1.15.0-shim'


# ----------------------------------------------------------------------------------------------
# dtypes
class DType:
    def __init__(self, name, tdt):
        self.name, self.t = name, tdt

    def __repr__(self):
        return 'tf.' + self.name

    def __eq__(self, o):
        return _dt(o).name == self.name if o is not None else False

    def __hash__(self):
        return hash(self.name)

    @property
    def as_numpy_dtype(self):
        return np.dtype(self.name if self.name != 'bool' else 'bool').type


float32 = DType('float32', torch.float32)
float64 = DType('float64', torch.float64)
int32 = DType('int32', torch.int32)
int64 = DType('int64', torch.int64)
uint8 = DType('uint8', torch.uint8)
bool = DType('bool', torch.bool)  # noqa: A001  (tf.bool)
_DTYPES = {d.name: d for d in (float32, float64, int32, int64, uint8, bool)}
_py_bool = _bi.bool
_py_slice = _bi.slice
_py_range, _py_abs, _py_pow = _bi.range, _bi.abs, _bi.pow


def _dt(x):
    if isinstance(x, DType):
        return x
    if isinstance(x, str):
        return _DTYPES[x]
    if isinstance(x, torch.dtype):
        return [d for d in _DTYPES.values() if d.t == x][0]
    return _DTYPES[np.dtype(x).name]


# ----------------------------------------------------------------------------------------------
# static shapes (only what the reference reads at graph-construction time)
class Dimension(int):
    """TF1's Dimension behaves like an int when known."""
    @property
    def value(self):
        return int(self)


class TensorShape:
    def __init__(self, dims):
        self.dims = None if dims is None else [None if d is None else Dimension(d) for d in dims]

    def __getitem__(self, i):
        if self.dims is None:
            raise ValueError('tfshim: static rank unknown')
        if isinstance(i, _py_slice):
            return TensorShape(self.dims[i])
        return self.dims[i]

    def __len__(self):
        return len(self.dims)

    def __iter__(self):
        return iter(self.dims)

    def as_list(self):
        return [None if d is None else int(d) for d in self.dims]

    @property
    def ndims(self):
        return None if self.dims is None else len(self.dims)

    def __repr__(self):
        return 'TensorShape(%s)' % (self.dims,)


def _sshape(x):
    """Static shape list (None entries for unknown dims) or None when the rank is unknown."""
    if isinstance(x, Tensor):
        return x._sshape
    if isinstance(x, (list, tuple)) and any(isinstance(e, Tensor) for e in x):
        return [len(x)]
    try:
        return list(np.shape(x))
    except Exception:
        return None


def _bcast(a, b):
    if a is None or b is None:
        return None
    n = max(len(a), len(b))
    a = [1] * (n - len(a)) + list(a)
    b = [1] * (n - len(b)) + list(b)
    out = []
    for x, y in zip(a, b):
        if x == 1:
            out.append(y)
        elif y == 1:
            out.append(x)
        elif x is None:
            out.append(y)
        else:
            out.append(x)
    return out


# ----------------------------------------------------------------------------------------------
# graph nodes
_NAME_STACK = []
_ALL_VARIABLES = []


def _conv_other(other, like):
    """TF converts python scalars / numpy values next to a Tensor to that Tensor's dtype."""
    if isinstance(other, torch.Tensor):
        return other
    return torch.as_tensor(np.asarray(other), dtype=like.dtype)


def _to_t(x, dtype=None):
    """Default tf.convert_to_tensor dtype inference: python float -> float32, int -> int32."""
    if isinstance(x, torch.Tensor):
        return x if dtype is None else x.to(_dt(dtype).t)
    if isinstance(x, (list, tuple)) and any(isinstance(e, torch.Tensor) for e in _flatten(x)):
        return torch.stack([_to_t(e, dtype) for e in x])
    a = np.asarray(x)
    if dtype is not None:
        return torch.as_tensor(a).to(_dt(dtype).t)
    if not isinstance(x, (np.ndarray, np.generic)):
        if a.dtype == np.float64:
            a = a.astype(np.float32)
        elif a.dtype == np.int64:
            a = a.astype(np.int32)
    return torch.as_tensor(a)


def _flatten(x):
    if isinstance(x, (list, tuple)):
        for e in x:
            yield from _flatten(e)
    else:
        yield x


def _ev(x, env):
    if isinstance(x, Tensor):
        k = id(x)
        if k in env:
            return env[k]
        v = x._compute(env)
        env[k] = v
        return v
    if isinstance(x, (list, tuple)):
        return [_ev(e, env) for e in x]
    return x


def _int(x):
    if isinstance(x, torch.Tensor):
        return int(x.item())
    return int(x)


def _ints(x):
    """Shape-like argument (tensor, list mixing ints / Dimensions / scalar tensors) -> [int]."""
    if isinstance(x, torch.Tensor):
        return [int(v) for v in x.reshape(-1).tolist()]
    if isinstance(x, (list, tuple)):
        out = []
        for e in x:
            if isinstance(e, (list, tuple)) or (isinstance(e, torch.Tensor) and e.dim() > 0) \
                    or (isinstance(e, np.ndarray) and e.ndim > 0):
                out += _ints(e)
            else:
                out.append(_int(e))
        return out
    if isinstance(x, np.ndarray):
        return [int(v) for v in x.reshape(-1)]
    return [_int(x)]


class Tensor:
    __array_ufunc__ = None
    __array_priority__ = 1000

    def __init__(self, op, inputs, fn, sshape=None, dtype=None, name=None):
        self.op_type, self.inputs, self.fn = op, inputs, fn
        self._sshape = sshape
        self._dtype = None if dtype is None else _dt(dtype)
        self.name = '/'.join(_NAME_STACK + [name or op])

    def _compute(self, env):
        return self.fn(*[_ev(i, env) for i in self.inputs])

    # ---- static info
    @property
    def shape(self):
        return TensorShape(self._sshape)

    def get_shape(self):
        return self.shape

    @property
    def dtype(self):
        return self._dtype

    def __repr__(self):
        return '<tfshim.Tensor %s %s>' % (self.name, self._sshape)

    def __hash__(self):
        return id(self)

    def __bool__(self):
        raise TypeError('tfshim: a symbolic Tensor has no truth value')

    def __iter__(self):
        raise TypeError('tfshim: Tensor is not iterable')

    # ---- operators
    def __add__(self, o): return add(self, o)
    def __radd__(self, o): return add(o, self)
    def __sub__(self, o): return subtract(self, o)
    def __rsub__(self, o): return subtract(o, self)
    def __mul__(self, o): return multiply(self, o)
    def __rmul__(self, o): return multiply(o, self)
    def __truediv__(self, o): return divide(self, o)
    def __rtruediv__(self, o): return divide(o, self)
    def __floordiv__(self, o): return _binary('FloorDiv', lambda a, b: torch.floor_divide(a, b), self, o)
    def __neg__(self): return _unary('Neg', torch.neg, self)
    def __pow__(self, o): return pow(self, o)
    def __rpow__(self, o): return pow(o, self)
    def __abs__(self): return abs(self)
    def __gt__(self, o): return greater(self, o)
    def __ge__(self, o): return _binary('GreaterEqual', torch.ge, self, o, dtype=bool)
    def __lt__(self, o): return _binary('Less', torch.lt, self, o, dtype=bool)
    def __le__(self, o): return _binary('LessEqual', torch.le, self, o, dtype=bool)
    def __eq__(self, o): return id(self) == id(o)      # TF1: identity, not elementwise

    def __getitem__(self, idx):
        return _getitem(self, idx)


def _guess_dtype(*xs):
    for x in xs:
        if isinstance(x, Tensor) and x._dtype is not None:
            return x._dtype
    return None


def _unary(op, f, x, sshape='same', dtype=None):
    ss = _sshape(x) if sshape == 'same' else sshape
    return Tensor(op, [x], lambda a: f(_to_t(a)), ss, dtype or _guess_dtype(x))


def _binary(op, f, a, b, dtype=None):
    def run(x, y):
        if isinstance(x, torch.Tensor) and not isinstance(y, torch.Tensor):
            y = _conv_other(y, x)
        elif isinstance(y, torch.Tensor) and not isinstance(x, torch.Tensor):
            x = _conv_other(x, y)
        elif not isinstance(x, torch.Tensor):
            x, y = _to_t(x), _to_t(y)
        return f(x, y)
    return Tensor(op, [a, b], run, _bcast(_sshape(a), _sshape(b)), dtype or _guess_dtype(a, b))


def _getitem(x, idx):
    if not isinstance(idx, tuple):
        idx = (idx,)

    def run(a, *dyn):
        dyn = list(dyn)
        a = _to_t(a)
        # expand Ellipsis
        n_spec = sum(1 for i in idx if i is not None and i is not Ellipsis)
        full = []
        for i in idx:
            if i is Ellipsis:
                full += [_py_slice(None)] * (a.dim() - n_spec)
            else:
                full.append(i)
        out, dim = a, 0
        for i in full:
            if i is None:
                out = out.unsqueeze(dim)
                dim += 1
            elif isinstance(i, _py_slice):
                start, stop, step = [(_int(dyn.pop(0)) if isinstance(v, Tensor) else v) for v in (i.start, i.stop, i.step)]
                if step is not None and step < 0:
                    n = out.shape[dim]
                    ids = list(_py_range(n))[_py_slice(start, stop, step)]
                    out = out.index_select(dim, torch.tensor(ids, dtype=torch.long))
                else:
                    sl = [_py_slice(None)] * out.dim()
                    sl[dim] = _py_slice(start, stop, step)
                    out = out[tuple(sl)]
                dim += 1
            else:
                k = _int(dyn.pop(0)) if isinstance(i, Tensor) else int(i)
                out = out.select(dim, k)
        return out

    dyn_in = []
    for i in idx:
        if isinstance(i, _py_slice):
            dyn_in += [v for v in (i.start, i.stop, i.step) if isinstance(v, Tensor)]
        elif isinstance(i, Tensor):
            dyn_in.append(i)
    # static shape
    ss = _sshape(x)
    out_ss = None
    if ss is not None:
        n_spec = sum(1 for i in idx if i is not None and i is not Ellipsis)
        full = []
        for i in idx:
            if i is Ellipsis:
                full += [_py_slice(None)] * (len(ss) - n_spec)
            else:
                full.append(i)
        out_ss, d = [], 0
        for i in full:
            if i is None:
                out_ss.append(1)
            elif isinstance(i, _py_slice):
                if i == _py_slice(None) or (i.start is None and i.stop is None):
                    out_ss.append(ss[d])
                elif ss[d] is not None and not any(isinstance(v, Tensor) for v in (i.start, i.stop, i.step)):
                    out_ss.append(len(_py_range(ss[d])[i]))
                else:
                    out_ss.append(None)
                d += 1
            else:
                d += 1
        out_ss += ss[d:]
    t = Tensor('StridedSlice', [x] + dyn_in, run, out_ss, _guess_dtype(x))
    if x.op_type == 'Shape' and len(idx) == 1 and isinstance(idx[0], int):
        t.shape_elem = (x.inputs[0], idx[0])
    return t


# ----------------------------------------------------------------------------------------------
# sources
def placeholder(dtype, shape=None, name=None):
    def missing():
        raise RuntimeError('tfshim: placeholder %s was not fed' % t.name)
    t = Tensor('Placeholder', [], missing, None if shape is None else list(shape), dtype, name)
    return t


def constant(value, dtype=None, shape=None, name=None):
    v = _to_t(value, dtype)
    if shape is not None:
        v = v.expand(*shape) if v.dim() == 0 else v.reshape(*shape)
    return Tensor('Const', [], lambda: v, list(v.shape), _dt(v.dtype), name)


def convert_to_tensor(value, dtype=None, name=None, preferred_dtype=None):
    if isinstance(value, Tensor):
        return value
    if isinstance(value, (list, tuple)) and any(isinstance(e, Tensor) for e in _flatten(value)):
        return stack(list(value))
    return constant(value, dtype)


class Variable(Tensor):
    def __init__(self, initial_value=None, trainable=True, validate_shape=True, name=None, dtype=None):
        ss = _sshape(initial_value) if validate_shape else None
        Tensor.__init__(self, 'VariableV2', [], None, ss, dtype or _guess_dtype(initial_value) or float32,
                        name or 'Variable')
        self.initial_value = initial_value
        _ALL_VARIABLES.append(self)

    def _compute(self, env):
        sess = env['__session__']
        if id(self) not in sess._vars:
            raise RuntimeError('tfshim: variable %s used before initialisation' % self.name)
        leaf = sess._vars[id(self)].detach().clone()
        if leaf.is_floating_point():
            leaf.requires_grad_(True)
        env.setdefault('__leaves__', {})[id(self)] = leaf
        return leaf

    @property
    def initializer(self):
        return _InitOp([self])


class Operation:
    def _run(self, sess, env):
        raise NotImplementedError


class _InitOp(Operation):
    def __init__(self, variables):
        self.variables = list(variables)

    def _run(self, sess, env):
        for v in self.variables:
            iv = v.initial_value
            if callable(iv) and not isinstance(iv, Tensor):
                val = iv(sess, env)
            else:
                val = _to_t(_ev(iv, env))
            sess._vars[id(v)] = val.detach().clone().to(v._dtype.t if v._dtype else val.dtype)


class _GroupOp(Operation):
    def __init__(self, ops):
        self.ops = ops

    def _run(self, sess, env):
        for o in self.ops:
            o._run(sess, env)


def variables_initializer(var_list, name='init'):
    return _InitOp(var_list)


def global_variables_initializer():
    return _InitOp(list(_ALL_VARIABLES))


def zeros_initializer():
    return 'zeros'


# ----------------------------------------------------------------------------------------------
# session
class Session:
    def __init__(self, graph=None, config=None):
        self._vars = {}
        self.graph = graph

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def close(self):
        pass

    def run(self, fetches, feed_dict=None):
        env = {'__session__': self}
        for k, v in (feed_dict or {}).items():
            if not isinstance(k, Tensor):
                raise TypeError('tfshim: feed key must be a Tensor, got %r' % (k,))
            if isinstance(v, (list, tuple)) and any(isinstance(e, torch.Tensor) for e in v):
                v = torch.stack(list(v))
            dt = k._dtype.t if k._dtype is not None else None
            t = v if isinstance(v, torch.Tensor) else torch.as_tensor(np.asarray(v))
            if dt is not None:
                t = t.to(dt)
            elif t.dtype == torch.float64:
                t = t.to(torch.float32)
            env[id(k)] = t
        ops = []

        def tensors_first(f):
            if isinstance(f, (list, tuple)):
                return [tensors_first(e) for e in f]
            if isinstance(f, dict):
                return {k: tensors_first(e) for k, e in f.items()}
            if isinstance(f, Operation):
                ops.append(f)
                return None
            if isinstance(f, Tensor):
                with torch.enable_grad():
                    v = _to_t(_ev(f, env))
                return v
            if isinstance(f, TopKV2):
                return TopKV2(tensors_first(f.values), tensors_first(f.indices))
            raise TypeError('tfshim: cannot fetch %r' % (f,))

        out = tensors_first(fetches)
        for o in ops:
            with torch.enable_grad():
                o._run(self, env)

        def to_np(v):
            if isinstance(v, list):
                return [to_np(e) for e in v]
            if isinstance(v, dict):
                return {k: to_np(e) for k, e in v.items()}
            if isinstance(v, TopKV2):
                return TopKV2(to_np(v.values), to_np(v.indices))
            if v is None:
                return None
            a = v.detach().numpy().copy()
            return a if a.ndim else a[()]
        return to_np(out)


InteractiveSession = Session


# ----------------------------------------------------------------------------------------------
# GraphDef import (inception5h, styler_base.py:17-31,53-57,91-94).  The protobuf messages are REAL protobuf
# messages (google.protobuf, declared on the fly with TensorFlow's field numbers), so ParseFromString and the
# reference's in-place attribute edit ``n.attr['strides'].list.i[1:3] = [1,1]`` run on the genuine library; the
# engine's own wire-format reader (lnst/graphdef.py) is not involved on this side.
_IMPORTED = {}          # 'import/<node>:0' -> Tensor   (one default graph, like the reference uses it)
_IMPORTED_OPS = []


def _graph_messages():
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    Fd = descriptor_pb2.FieldDescriptorProto
    fd = descriptor_pb2.FileDescriptorProto(name='tfshim_graph.proto', package='tfshim', syntax='proto3')
    O, R = Fd.LABEL_OPTIONAL, Fd.LABEL_REPEATED

    def fields(m, spec):
        for fname, num, typ, label, tname in spec:
            f = m.field.add(name=fname, number=num, type=typ, label=label)
            if tname:
                f.type_name = '.tfshim.' + tname

    m = fd.message_type.add(name='TensorShapeProto')
    d = m.nested_type.add(name='Dim')
    fields(d, [('size', 1, Fd.TYPE_INT64, O, None), ('name', 2, Fd.TYPE_STRING, O, None)])
    fields(m, [('dim', 2, Fd.TYPE_MESSAGE, R, 'TensorShapeProto.Dim'), ('unknown_rank', 3, Fd.TYPE_BOOL, O, None)])
    m = fd.message_type.add(name='TensorProto')
    fields(m, [('dtype', 1, Fd.TYPE_INT32, O, None), ('tensor_shape', 2, Fd.TYPE_MESSAGE, O, 'TensorShapeProto'),
               ('version_number', 3, Fd.TYPE_INT32, O, None), ('tensor_content', 4, Fd.TYPE_BYTES, O, None),
               ('float_val', 5, Fd.TYPE_FLOAT, R, None), ('double_val', 6, Fd.TYPE_DOUBLE, R, None),
               ('int_val', 7, Fd.TYPE_INT32, R, None), ('string_val', 8, Fd.TYPE_BYTES, R, None)])
    m = fd.message_type.add(name='AttrValue')
    lv = m.nested_type.add(name='ListValue')
    fields(lv, [('s', 2, Fd.TYPE_BYTES, R, None), ('i', 3, Fd.TYPE_INT64, R, None), ('f', 4, Fd.TYPE_FLOAT, R, None),
                ('b', 5, Fd.TYPE_BOOL, R, None), ('type', 6, Fd.TYPE_INT32, R, None)])
    fields(m, [('list', 1, Fd.TYPE_MESSAGE, O, 'AttrValue.ListValue'), ('s', 2, Fd.TYPE_BYTES, O, None),
               ('i', 3, Fd.TYPE_INT64, O, None), ('f', 4, Fd.TYPE_FLOAT, O, None), ('b', 5, Fd.TYPE_BOOL, O, None),
               ('type', 6, Fd.TYPE_INT32, O, None), ('shape', 7, Fd.TYPE_MESSAGE, O, 'TensorShapeProto'),
               ('tensor', 8, Fd.TYPE_MESSAGE, O, 'TensorProto')])
    m = fd.message_type.add(name='NodeDef')
    e = m.nested_type.add(name='AttrEntry')
    e.options.map_entry = True
    fields(e, [('key', 1, Fd.TYPE_STRING, O, None), ('value', 2, Fd.TYPE_MESSAGE, O, 'AttrValue')])
    fields(m, [('name', 1, Fd.TYPE_STRING, O, None), ('op', 2, Fd.TYPE_STRING, O, None),
               ('input', 3, Fd.TYPE_STRING, R, None), ('device', 4, Fd.TYPE_STRING, O, None),
               ('attr', 5, Fd.TYPE_MESSAGE, R, 'NodeDef.AttrEntry')])
    m = fd.message_type.add(name='GraphDef')
    fields(m, [('node', 1, Fd.TYPE_MESSAGE, R, 'NodeDef')])
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    get = getattr(message_factory, 'GetMessageClass', None)
    if get is None:                                            # protobuf < 4.21
        get = message_factory.MessageFactory(pool).GetPrototype
    return get(pool.FindMessageTypeByName('tfshim.GraphDef'))


_GRAPHDEF_CLS = None


def GraphDef():
    global _GRAPHDEF_CLS
    if _GRAPHDEF_CLS is None:
        _GRAPHDEF_CLS = _graph_messages()
    return _GRAPHDEF_CLS()


def _tensor_proto_to_np(t):
    shape = [int(d.size) for d in t.tensor_shape.dim]
    if t.dtype == 1:
        dt, vals = np.float32, list(t.float_val)
    elif t.dtype == 3:
        dt, vals = np.int32, list(t.int_val)
    else:
        raise NotImplementedError('tfshim: Const dtype %d' % t.dtype)
    if t.tensor_content:
        arr = np.frombuffer(t.tensor_content, dt)
    else:
        n = int(np.prod(shape)) if shape else 1
        arr = np.asarray(vals, dt)
        if arr.size == 1 and n > 1:
            arr = np.full(n, arr[0], dt)
    return arr.reshape(shape).copy()


def _lrn(x, depth_radius=5, bias=1.0, alpha=1.0, beta=0.5, name=None):
    """tf.nn.local_response_normalization (core/kernels/lrn_op.cc): sqr_sum over [c-r, c+r], alpha NOT divided by n."""
    r = int(depth_radius)

    def f(a):
        sq = F.pad(a * a, (r, r))
        s = sq[..., 0:a.shape[-1]]
        for i in _py_range(1, 2 * r + 1):
            s = s + sq[..., i:i + a.shape[-1]]
        return a * torch.pow(bias + alpha * s, -beta)
    return _unary('LRN', f, x)


class _ImportedOp:
    def __init__(self, name, type_):
        self.name, self.type = name, type_


def import_graph_def(graph_def, input_map=None, return_elements=None, name=None):
    """tf.import_graph_def for frozen inference graphs: every node becomes a lazy Tensor named
    '<prefix>/<node>:0' (prefix 'import'), placeholders are replaced through ``input_map``."""
    prefix = 'import' if name is None else name
    input_map = {k.split(':')[0]: v for k, v in (input_map or {}).items()}
    local = {}

    def ref(n):
        n = n.split(':')[0]
        return local[n]

    for nd in graph_def.node:
        a = nd.attr
        ins = [i for i in nd.input if not i.startswith('^')]
        if nd.name in input_map:
            t = input_map[nd.name]
        elif nd.op == 'Placeholder':
            t = placeholder(float32, name=nd.name)
        elif nd.op == 'Const':
            t = constant(_tensor_proto_to_np(a['value'].tensor))
        elif nd.op == 'Conv2D':
            t = nn.conv2d(ref(ins[0]), ref(ins[1]), list(a['strides'].list.i), a['padding'].s.decode())
        elif nd.op == 'BiasAdd':
            t = add(ref(ins[0]), ref(ins[1]))
        elif nd.op == 'Relu':
            t = nn.relu(ref(ins[0]))
        elif nd.op == 'MaxPool':
            t = nn.max_pool(ref(ins[0]), list(a['ksize'].list.i), list(a['strides'].list.i), a['padding'].s.decode())
        elif nd.op == 'LRN':
            kw = {}
            for k in ('depth_radius', 'bias', 'alpha', 'beta'):
                if k in a:
                    kw[k] = a[k].i if k == 'depth_radius' else a[k].f
            t = _lrn(ref(ins[0]), **kw)
        elif nd.op == 'Concat':
            t = concat([ref(i) for i in ins[1:]], int(_tensor_value(ref(ins[0]))))
        elif nd.op == 'ConcatV2':
            t = concat([ref(i) for i in ins[:-1]], int(_tensor_value(ref(ins[-1]))))
        elif nd.op == 'Identity':
            t = identity(ref(ins[0]))
        elif nd.op == 'AvgPool':
            t = nn.avg_pool(ref(ins[0]), list(a['ksize'].list.i), list(a['strides'].list.i), a['padding'].s.decode())
        elif nd.op == 'Reshape':
            t = reshape(ref(ins[0]), [int(v) for v in np.asarray(_tensor_value_np(ref(ins[1]))).reshape(-1)])
        elif nd.op == 'MatMul':
            t = matmul(ref(ins[0]), ref(ins[1]), transpose_a=a['transpose_a'].b if 'transpose_a' in a else False,
                       transpose_b=a['transpose_b'].b if 'transpose_b' in a else False)
        else:
            raise NotImplementedError('tfshim.import_graph_def: op %s (%s)' % (nd.op, nd.name))
        local[nd.name] = t
        _IMPORTED['%s/%s:0' % (prefix, nd.name)] = t
        _IMPORTED_OPS.append(_ImportedOp('%s/%s' % (prefix, nd.name), nd.op))
    if return_elements:
        return [ref(e) for e in return_elements]


def _tensor_value(t):
    return _to_t(_ev(t, {})).item() if isinstance(t, Tensor) else t


def _tensor_value_np(t):
    return _to_t(_ev(t, {})).numpy() if isinstance(t, Tensor) else np.asarray(t)


class Graph:
    def __init__(self):                      # a fresh graph per Styler (styler_base.py:20): forget earlier imports
        _IMPORTED.clear()
        del _IMPORTED_OPS[:]

    def get_operations(self):
        return list(_IMPORTED_OPS)

    def get_tensor_by_name(self, name):
        return _IMPORTED[name]


@contextlib.contextmanager
def name_scope(name, default_name=None, values=None):
    _NAME_STACK.append(name or default_name)
    try:
        yield '/'.join(_NAME_STACK)
    finally:
        _NAME_STACK.pop()


@contextlib.contextmanager
def variable_scope(name_or_scope, default_name=None, values=None, reuse=None):
    _NAME_STACK.append(name_or_scope or default_name)
    try:
        yield '/'.join(_NAME_STACK)
    finally:
        _NAME_STACK.pop()


# ----------------------------------------------------------------------------------------------
# elementwise math
def add(a, b, name=None): return _binary('Add', torch.add, a, b)
def subtract(a, b, name=None): return _binary('Sub', torch.sub, a, b)
def multiply(a, b, name=None): return _binary('Mul', torch.mul, a, b)


def divide(a, b, name=None):
    # python3 '/' on TF tensors is RealDiv for floats (ints would be cast to float64: not used on this path)
    return _binary('RealDiv', torch.true_divide, a, b)


def pow(a, b, name=None):  # noqa: A001
    return _binary('Pow', torch.pow, a, b)


def maximum(a, b, name=None):
    # gradient to x where x >= y, else to y (math_grad.py _MaximumMinimumGrad)
    def f(x, y):
        x, y = torch.broadcast_tensors(x, y)
        return torch.where(x >= y, x, y)
    return _binary('Maximum', f, a, b)


def minimum(a, b, name=None):
    def f(x, y):
        x, y = torch.broadcast_tensors(x, y)
        return torch.where(x <= y, x, y)
    return _binary('Minimum', f, a, b)


def clip_by_value(t, clip_value_min, clip_value_max, name=None):
    # clip_ops.py (1.15): t_min = minimum(values, max); t_max = maximum(t_min, min)
    return maximum(minimum(t, clip_value_max), clip_value_min)


def sqrt(x, name=None): return _unary('Sqrt', torch.sqrt, x)
def square(x, name=None): return _unary('Square', lambda a: a * a, x)
def exp(x, name=None): return _unary('Exp', torch.exp, x)
def log(x, name=None): return _unary('Log', torch.log, x)
def abs(x, name=None): return _unary('Abs', torch.abs, x)  # noqa: A001
def floor(x, name=None): return _unary('Floor', torch.floor, x)
def ceil(x, name=None): return _unary('Ceil', torch.ceil, x)
def zeros_like(x, dtype=None, name=None): return _unary('ZerosLike', torch.zeros_like, x)
def ones_like(x, dtype=None, name=None): return _unary('OnesLike', torch.ones_like, x)
def identity(x, name=None): return _unary('Identity', lambda a: a, x)
def stop_gradient(x, name=None): return _unary('StopGradient', lambda a: a.detach(), x)


def squared_difference(a, b, name=None):
    return _binary('SquaredDifference', lambda x, y: (x - y) * (x - y), a, b)


def mod(a, b, name=None): return _binary('FloorMod', torch.remainder, a, b)
def equal(a, b, name=None): return _binary('Equal', torch.eq, a, b, dtype=bool)
def not_equal(a, b, name=None): return _binary('NotEqual', torch.ne, a, b, dtype=bool)
def greater(a, b, name=None): return _binary('Greater', torch.gt, a, b, dtype=bool)
def less(a, b, name=None): return _binary('Less', torch.lt, a, b, dtype=bool)
def logical_and(a, b, name=None): return _binary('LogicalAnd', torch.logical_and, a, b, dtype=bool)
def logical_or(a, b, name=None): return _binary('LogicalOr', torch.logical_or, a, b, dtype=bool)


def cast(x, dtype, name=None):
    d = _dt(dtype)

    def f(a):
        a = _to_t(a)
        if a.is_floating_point() and not d.t.is_floating_point and d.t != torch.bool:
            return torch.trunc(a).to(d.t)          # C++ float->int conversion truncates
        return a.to(d.t)
    return Tensor('Cast', [x], f, _sshape(x), d)


def to_int32(x, name=None): return cast(x, int32)
def to_float(x, name=None): return cast(x, float32)


def where(condition, x=None, y=None, name=None):
    if x is None:
        raise NotImplementedError('tfshim: single-argument tf.where')

    def f(c, a, b):
        a = _to_t(a) if not isinstance(a, torch.Tensor) else a
        b = _conv_other(b, a)
        if c.dim() == 1 and a.dim() > 1:           # TF1 where: vector condition selects rows
            c = c.reshape([-1] + [1] * (a.dim() - 1))
        return torch.where(c, a, b)
    return Tensor('Select', [condition, x, y], f, _sshape(x), _guess_dtype(x, y))


def add_n(inputs, name=None):
    def f(xs):
        out = xs[0]
        for x in xs[1:]:
            out = out + x
        return out
    return Tensor('AddN', [list(inputs)], f, _sshape(inputs[0]), _guess_dtype(*inputs))


# ----------------------------------------------------------------------------------------------
# shapes and layout
def shape(x, name=None, out_type=None):
    t = Tensor('Shape', [x], lambda a: torch.tensor(list(_to_t(a).shape), dtype=torch.int32),
               None if _sshape(x) is None else [len(_sshape(x))], int32)
    return t


def _static_from_shape_arg(shp):
    if isinstance(shp, Tensor):
        if shp.op_type == 'Shape':
            return _sshape(shp.inputs[0])
        return None
    out = []
    for e in shp:
        if isinstance(e, Tensor):
            se = getattr(e, 'shape_elem', None)
            if se is not None and _sshape(se[0]) is not None:
                out.append(_sshape(se[0])[se[1]])
            else:
                out.append(None)
        elif isinstance(e, (list, tuple, np.ndarray)):
            return None
        else:
            out.append(None if int(e) == -1 else int(e))
    return out


def reshape(tensor, shape, name=None):  # noqa: A002
    return Tensor('Reshape', [tensor, shape], lambda a, s: _to_t(a).reshape(_ints(s)),
                  _static_from_shape_arg(shape), _guess_dtype(tensor))


def expand_dims(x, axis=None, name=None, dim=None):
    axis = dim if axis is None else axis
    ss = _sshape(x)
    if ss is not None:
        ss = list(ss)
        ss.insert(axis if axis >= 0 else len(ss) + 1 + axis, 1)
    return Tensor('ExpandDims', [x], lambda a: _to_t(a).unsqueeze(axis), ss, _guess_dtype(x))


def squeeze(x, axis=None, name=None):
    return Tensor('Squeeze', [x], lambda a: a.squeeze() if axis is None else a.squeeze(axis), None, _guess_dtype(x))


def concat(values, axis, name=None):
    def f(vs):
        ts = []
        ref = next((v for v in vs if isinstance(v, torch.Tensor)), None)
        for v in vs:
            if isinstance(v, torch.Tensor):
                ts.append(v)
            elif isinstance(v, (list, tuple)):
                t = torch.stack([e if isinstance(e, torch.Tensor) else torch.as_tensor(_int(e)) for e in v]) \
                    if len(v) else torch.zeros(0)
                ts.append(t.to(ref.dtype) if ref is not None else t.to(torch.int32))
            else:
                ts.append(_conv_other(v, ref) if ref is not None else _to_t(v))
        return torch.cat(ts, dim=axis)
    ss = None
    s0 = [_sshape(v) for v in values]
    if all(s is not None for s in s0) and len({len(s) for s in s0}) == 1:
        ss = list(s0[0])
        ax = axis if axis >= 0 else len(ss) + axis
        ss[ax] = None if any(s[ax] is None for s in s0) else sum(s[ax] for s in s0)
        for d in _py_range(len(ss)):
            if d != ax and ss[d] is None:
                known = [s[d] for s in s0 if s[d] is not None]
                ss[d] = known[0] if known else None
    return Tensor('ConcatV2', [list(values)], f, ss, _guess_dtype(*[v for v in values if isinstance(v, Tensor)]))


def stack(values, axis=0, name=None):
    def f(vs):
        ref = next((v for v in vs if isinstance(v, torch.Tensor)), None)
        ts = [v if isinstance(v, torch.Tensor) else (_conv_other(v, ref) if ref is not None else _to_t(v)) for v in vs]
        return torch.stack(ts, dim=axis)
    ss = _sshape(values[0]) if len(values) else None
    if ss is not None:
        ss = list(ss)
        ss.insert(axis if axis >= 0 else len(ss) + 1 + axis, len(values))
    return Tensor('Pack', [list(values)], f, ss, _guess_dtype(*[v for v in values if isinstance(v, Tensor)]))


def tile(x, multiples, name=None):
    return Tensor('Tile', [x, multiples], lambda a, m: _to_t(a).repeat(*_ints(m)), None, _guess_dtype(x))


def transpose(a, perm=None, name=None):
    def f(x):
        return x.permute(*perm) if perm is not None else x.permute(*reversed(_py_range(x.dim())))
    ss = _sshape(a)
    if ss is not None:
        ss = [ss[p] for p in perm] if perm is not None else ss[::-1]
    return Tensor('Transpose', [a], f, ss, _guess_dtype(a))


def slice(input_, begin, size, name=None):  # noqa: A001
    def f(a, b, s):
        b, s = _ints(b), _ints(s)
        idx = tuple(_py_slice(bi, None if si == -1 else bi + si) for bi, si in zip(b, s))
        return a[idx]
    return Tensor('Slice', [input_, begin, size], f, None, _guess_dtype(input_))




def zeros(shape, dtype=float32, name=None):  # noqa: A002
    return Tensor('Zeros', [shape], lambda s: torch.zeros(_ints(s) if not (isinstance(s, list) and len(s) == 0) else [],
                                                           dtype=_dt(dtype).t), None, dtype)


def ones(shape, dtype=float32, name=None):  # noqa: A002
    return Tensor('Ones', [shape], lambda s: torch.ones(_ints(s), dtype=_dt(dtype).t), None, dtype)


def range(start, limit=None, delta=1, dtype=None, name=None):  # noqa: A001
    if limit is None:
        start, limit = 0, start

    def f(a, b, d):
        return torch.arange(_int(a), _int(b), _int(d), dtype=torch.int32 if dtype is None else _dt(dtype).t)
    return Tensor('Range', [start, limit, delta], f, None, dtype or int32)


def linspace(start, stop, num, name=None):
    # sequence_ops.cc LinSpaceOp (1.15): step = (stop-start)/(num-1); out[i] = start + step*i, all in T
    def f(a, b, n):
        a, b, n = _to_t(a).to(torch.float32), _to_t(b).to(torch.float32), _int(n)
        if n == 1:
            return a.reshape(1)
        step = (b - a) / torch.tensor(float(n - 1), dtype=torch.float32)
        return a + step * torch.arange(n, dtype=torch.float32)
    return Tensor('LinSpace', [start, stop, num], f, None, float32)


def meshgrid(*args, **kwargs):
    indexing = kwargs.get('indexing', 'xy')
    n = len(args)
    return [Tensor('Meshgrid%d' % i, [list(args)],
                   (lambda xs, i=i: torch.meshgrid(*xs, indexing=indexing)[i].contiguous()), None, _guess_dtype(*args))
            for i in _py_range(n)]


def gather(params, indices, axis=0, name=None):
    def f(p, i):
        i = _to_t(i).long()
        return p.index_select(axis, i.reshape(-1)).reshape(list(p.shape[:axis]) + list(i.shape) + list(p.shape[axis + 1:]))
    return Tensor('GatherV2', [params, indices], f, None, _guess_dtype(params))


def gather_nd(params, indices, name=None):
    def f(p, i):
        i = _to_t(i).long()
        return p[tuple(i[..., k] for k in _py_range(i.shape[-1]))]
    return Tensor('GatherNd', [params, indices], f, None, _guess_dtype(params))


def boolean_mask(tensor, mask, name=None, axis=None):
    def f(t, m):
        t = _to_t(t)
        return t[m]
    return Tensor('BooleanMask', [tensor, mask], f, None, _guess_dtype(tensor))


def scatter_nd(indices, updates, shape, name=None):  # noqa: A002
    def f(idx, upd, shp):
        shp = _ints(shp)
        idx = _to_t(idx).long()
        upd = _to_t(upd)
        k = idx.shape[-1]
        idx = idx.reshape(-1, k)
        inner = shp[k:]
        upd = upd.reshape([idx.shape[0]] + inner)
        dims = torch.tensor(shp[:k], dtype=torch.long)
        ok = ((idx >= 0) & (idx < dims)).all(dim=1)         # GPU kernel: out-of-range updates are dropped
        strides = [int(np.prod(shp[j + 1:k])) for j in _py_range(k)]
        flat = (idx * torch.tensor(strides, dtype=torch.long)).sum(dim=1)
        if not _py_bool(ok.all()):
            flat, upd = flat[ok], upd[ok]
        out = torch.zeros([int(np.prod(shp[:k]))] + inner, dtype=upd.dtype)
        out = out.index_add(0, flat, upd)
        return out.reshape(shp)
    return Tensor('ScatterNd', [indices, updates, shape], f, None, _guess_dtype(updates))


def matmul(a, b, transpose_a=False, transpose_b=False, name=None):
    def f(x, y):
        if transpose_a:
            x = x.transpose(-1, -2)
        if transpose_b:
            y = y.transpose(-1, -2)
        if not x.is_floating_point():
            return (x.unsqueeze(-1) * y.unsqueeze(-3)).sum(dim=-2).to(x.dtype)
        return torch.matmul(x, y)
    sa, sb = _sshape(a), _sshape(b)
    ss = None
    if sa is not None and sb is not None and len(sa) == 2 and len(sb) == 2:
        ss = [sa[1] if transpose_a else sa[0], sb[0] if transpose_b else sb[1]]
    return Tensor('MatMul', [a, b], f, ss, _guess_dtype(a, b))


# ----------------------------------------------------------------------------------------------
# reductions
def _axes(axis, nd):
    if axis is None:
        return list(_py_range(nd))
    if isinstance(axis, (list, tuple)):
        return [a % nd for a in axis]
    return [axis % nd]


def _reduce(op, f):
    def red(input_tensor, axis=None, keepdims=False, name=None, keep_dims=None, reduction_indices=None):
        if keep_dims is not None:
            keepdims = keep_dims
        if reduction_indices is not None:
            axis = reduction_indices

        def run(a):
            a = _to_t(a)
            ax = _axes(axis, a.dim())
            if not ax:
                return a
            return f(a, ax, keepdims)
        return Tensor(op, [input_tensor], run, None, _guess_dtype(input_tensor))
    return red


reduce_sum = _reduce('Sum', lambda a, ax, k: a.sum(dim=ax, keepdim=k))
reduce_mean = _reduce('Mean', lambda a, ax, k: a.mean(dim=ax, keepdim=k))
reduce_all = _reduce('All', lambda a, ax, k: (a.to(torch.int32).sum(dim=ax, keepdim=k) == int(np.prod([a.shape[d] for d in ax]))))
reduce_any = _reduce('Any', lambda a, ax, k: a.to(torch.int32).sum(dim=ax, keepdim=k) > 0)


class _MinMaxTies(torch.autograd.Function):
    """math_grad.py _MinOrMaxGrad: grad * equal(y, x) / sum(equal(y, x))."""
    @staticmethod
    def forward(ctx, a, ax, keep, is_max):
        y = a.amax(dim=ax, keepdim=True) if is_max else a.amin(dim=ax, keepdim=True)
        ctx.save_for_backward(a, y)
        ctx.ax, ctx.keep = ax, keep
        return y if keep else y.squeeze(ax) if len(ax) < a.dim() else y.reshape([])

    @staticmethod
    def backward(ctx, g):
        a, y = ctx.saved_tensors
        ind = (a == y).to(a.dtype)
        num = ind.sum(dim=ctx.ax, keepdim=True)
        g = g.reshape(y.shape)
        return ind / num * g, None, None, None


reduce_max = _reduce('Max', lambda a, ax, k: _MinMaxTies.apply(a, ax, k, True))
reduce_min = _reduce('Min', lambda a, ax, k: _MinMaxTies.apply(a, ax, k, False))


def cumsum(x, axis=0, exclusive=False, reverse=False, name=None):
    def f(a):
        if reverse:
            a = a.flip(axis)
        c = a.cumsum(dim=axis)
        if exclusive:
            c = c - a
        return c.flip(axis) if reverse else c
    return _unary('Cumsum', f, x)


def argmin(x, axis=None, name=None, output_type=int64):
    return Tensor('ArgMin', [x], lambda a: a.argmin(dim=axis), None, output_type)


def argmax(x, axis=None, name=None, output_type=int64):
    return Tensor('ArgMax', [x], lambda a: a.argmax(dim=axis), None, output_type)


class TopKV2:
    def __init__(self, values, indices):
        self.values, self.indices = values, indices


def histogram_fixed_width(*a, **k):
    raise NotImplementedError('tfshim: histogram loss is outside the pinned path')


def py_func(*a, **k):
    raise NotImplementedError('tfshim: py_func')


def map_fn(*a, **k):
    raise NotImplementedError('tfshim: map_fn')


# ----------------------------------------------------------------------------------------------
# tf.nn
def _same_pad(n, k, s):
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return total // 2, total - total // 2


def _conv_nd(nd):
    def conv(input, filter=None, strides=None, padding=None, name=None, filters=None, **kw):  # noqa: A002
        w_in = filters if filter is None else filter

        def f(x, w):
            x = _to_t(x)
            w = _conv_other(w, x)
            perm_in = [0, nd + 1] + list(_py_range(1, nd + 1))
            xc = x.permute(*perm_in)
            wc = w.permute(*([nd + 1, nd] + list(_py_range(nd))))          # [k.., I, O] -> [O, I, k..]
            st = list(strides)[1:-1] if not isinstance(strides, int) else [strides] * nd
            if padding == 'SAME':
                pads = []
                for d in reversed(_py_range(nd)):
                    lo, hi = _same_pad(xc.shape[2 + d], wc.shape[2 + d], st[d])
                    pads += [lo, hi]
                xc = F.pad(xc, pads)
            y = (F.conv2d if nd == 2 else F.conv3d)(xc, wc, stride=st)
            return y.permute(*([0] + list(_py_range(2, nd + 2)) + [1]))
        ss = _sshape(input)
        ws = _sshape(w_in)
        out_ss = None
        if ss is not None:
            out_ss = list(ss)
            out_ss[-1] = ws[-1] if ws is not None else None
            if strides is not None and any(int(s) != 1 for s in (strides if not isinstance(strides, int) else [strides])):
                out_ss[1:-1] = [None] * nd
        return Tensor('Conv%dD' % nd, [input, w_in], f, out_ss, _guess_dtype(input))
    return conv


def _pool(kind, nd):
    def pool(value, ksize, strides, padding, name=None, data_format=None):
        def f(x):
            perm_in = [0, nd + 1] + list(_py_range(1, nd + 1))
            xc = x.permute(*perm_in)
            ks = list(ksize)[1:-1]
            st = list(strides)[1:-1]
            if padding == 'SAME':
                pads = []
                for d in reversed(_py_range(nd)):
                    lo, hi = _same_pad(xc.shape[2 + d], ks[d], st[d])
                    pads += [lo, hi]
                if kind == 'max':
                    xc = F.pad(xc, pads, value=float('-inf'))
                elif any(pads):
                    raise NotImplementedError('tfshim: SAME avg_pool with padding')
            fn = {('max', 2): F.max_pool2d, ('max', 3): F.max_pool3d, ('avg', 2): F.avg_pool2d, ('avg', 3): F.avg_pool3d}[(kind, nd)]
            y = fn(xc, ks, st)
            return y.permute(*([0] + list(_py_range(2, nd + 2)) + [1]))
        ss = _sshape(value)
        out_ss = None if ss is None else [ss[0]] + [None] * nd + [ss[-1]]
        return Tensor('%sPool%dD' % (kind, nd), [value], f, out_ss, _guess_dtype(value))
    return pool


def _top_k(input, k=1, sorted=True, name=None):  # noqa: A002
    def vals(a):
        return torch.topk(_to_t(a), k, dim=-1)[0]

    def inds(a):
        return torch.topk(_to_t(a), k, dim=-1)[1].to(torch.int32)
    return TopKV2(Tensor('TopKV2', [input], vals), Tensor('TopKV2i', [input], inds, dtype=int32))


nn = types.SimpleNamespace(
    relu=lambda x, name=None: maximum(x, 0.0) if False else _unary('Relu', torch.relu, x),
    conv2d=_conv_nd(2), conv3d=_conv_nd(3),
    max_pool=_pool('max', 2), max_pool3d=_pool('max', 3), avg_pool=_pool('avg', 2), avg_pool3d=_pool('avg', 3),
    top_k=_top_k,
    conv2d_transpose=None, conv3d_transpose=None,
    lrn=lambda *a, **k: _lrn(*a, **k), local_response_normalization=lambda *a, **k: _lrn(*a, **k),
    bias_add=lambda v, b, name=None: add(v, b),
)


# ----------------------------------------------------------------------------------------------
# tf.image (v1 resize: align_corners=False, legacy pixel mapping)
class _ResizeMethod:
    BILINEAR, NEAREST_NEIGHBOR, BICUBIC, AREA = 0, 1, 2, 3


def _resize_bilinear_1d(x, dim, out_n):
    """resize_bilinear_op.cc compute_interpolation_weights (legacy scaler: in = out * scale)."""
    in_n = x.shape[dim]
    scale = np.float32(in_n) / np.float32(out_n)
    o = np.arange(out_n, dtype=np.float32)
    src = o * scale                                           # float32 like the kernel
    lo = np.floor(src)
    hi = np.minimum(np.ceil(src), in_n - 1)
    lerp = torch.as_tensor((src - lo).astype(np.float32))
    shp = [1] * x.dim()
    shp[dim] = out_n
    lerp = lerp.reshape(shp)
    a = x.index_select(dim, torch.as_tensor(lo.astype(np.int64)))
    b = x.index_select(dim, torch.as_tensor(hi.astype(np.int64)))
    return a + (b - a) * lerp


def _bicubic_table():
    """resize_bicubic_op.cc InitCoeffsTable(A = -0.75), kTableSize = 1024."""
    n = 1 << 10
    a = -0.75
    tab = np.zeros((n + 1) * 2, dtype=np.float32)
    for i in _py_range(n + 1):
        x = np.float32(i * 1.0 / n)
        tab[i * 2] = ((a + 2) * x - (a + 3)) * x * x + 1
        x = np.float32(x + 1.0)
        tab[i * 2 + 1] = ((a * x - 5 * a) * x + 8 * a) * x - 4 * a
    return tab


_BICUBIC = None


def _resize_bicubic_1d(x, dim, out_n):
    """resize_bicubic_op.cc GetWeightsAndIndices<LegacyScaler, use_keys_cubic=false>."""
    global _BICUBIC
    if _BICUBIC is None:
        _BICUBIC = _bicubic_table()
    n = 1 << 10
    in_n = x.shape[dim]
    scale = np.float32(in_n) / np.float32(out_n)
    o = np.arange(out_n, dtype=np.float32)
    src = o * scale
    fl = np.floor(src)
    delta = src - fl
    off = np.rint(delta * n).astype(np.int64)                # lrint
    w = np.stack([_BICUBIC[off * 2 + 1], _BICUBIC[off * 2], _BICUBIC[(n - off) * 2], _BICUBIC[(n - off) * 2 + 1]])
    idx = np.stack([fl - 1, fl, fl + 1, fl + 2]).astype(np.int64)
    idx = np.clip(idx, 0, in_n - 1)                           # Bound()
    out = 0
    shp = [1] * x.dim()
    shp[dim] = out_n
    for k in _py_range(4):
        out = out + x.index_select(dim, torch.as_tensor(idx[k])) * torch.as_tensor(w[k]).reshape(shp)
    return out


def _image_resize(images, size, method=_ResizeMethod.BILINEAR, align_corners=False, preserve_aspect_ratio=False, name=None):
    if align_corners:
        raise NotImplementedError
    fn = {_ResizeMethod.BILINEAR: _resize_bilinear_1d, _ResizeMethod.BICUBIC: _resize_bicubic_1d}[method]

    def f(x, s):
        x = _to_t(x)
        h, w = _ints(s)
        if h == x.shape[1] and w == x.shape[2]:
            return x                                         # resize_images_v1 returns the input unchanged
        if method == _ResizeMethod.BILINEAR:
            return fn(fn(x, 1, h), 2, w)                     # kernel: top/bottom lerp in x first, then y
        return fn(fn(x, 2, w), 1, h)
    ss = _sshape(images)
    out_ss = [None] * 4 if ss is None else [ss[0], None, None, ss[3]]       # resize_images: always rank 4
    return Tensor('Resize', [images, size], f, out_ss, _guess_dtype(images))


def _total_variation(images, name=None):
    def f(x):
        dh = (x[:, 1:] - x[:, :-1]).abs().sum(dim=[1, 2, 3])
        dw = (x[:, :, 1:] - x[:, :, :-1]).abs().sum(dim=[1, 2, 3])
        return dh + dw
    return _unary('TotalVariation', f, images, sshape=None)


image = types.SimpleNamespace(ResizeMethod=_ResizeMethod, resize=_image_resize, resize_images=_image_resize,
                              total_variation=_total_variation)


# ----------------------------------------------------------------------------------------------
# tf.train.AdamOptimizer
class AdamOptimizer:
    def __init__(self, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8, use_locking=False, name='Adam'):
        self._lr, self._b1, self._b2, self._eps = learning_rate, beta1, beta2, epsilon
        self._slots = {}
        self._b1p = self._b2p = None
        self._vars = []

    def minimize(self, loss, global_step=None, var_list=None, **kw):
        assert var_list is not None
        self._vars = list(var_list)
        for v in self._vars:
            def zeros_like_primary(sess, env, v=v):
                return torch.zeros_like(sess._vars[id(v)])      # slot_creator.create_zeros_slot (dynamic shape)
            self._slots[id(v)] = (Variable(zeros_like_primary, validate_shape=False, name=v.name + '/Adam'),
                                  Variable(zeros_like_primary, validate_shape=False, name=v.name + '/Adam_1'))
        self._b1p = Variable(np.float32(self._b1), name='beta1_power')
        self._b2p = Variable(np.float32(self._b2), name='beta2_power')
        return _AdamStep(self, loss)

    def variables(self):
        out = [self._b1p, self._b2p]
        for v in self._vars:
            out += list(self._slots[id(v)])
        return out


class _AdamStep(Operation):
    def __init__(self, opt, loss):
        self.opt, self.loss = opt, loss

    def _run(self, sess, env):
        o = self.opt
        loss = _ev(self.loss, env)
        leaves = env.get('__leaves__', {})
        used = [v for v in o._vars if id(v) in leaves]
        grads = torch.autograd.grad(loss, [leaves[id(v)] for v in used], allow_unused=True, retain_graph=True)
        f32 = torch.float32
        lr = _to_t(_ev(o._lr, env)).to(f32)
        b1p, b2p = sess._vars[id(o._b1p)].to(f32), sess._vars[id(o._b2p)].to(f32)
        b1, b2, eps = torch.tensor(o._b1, dtype=f32), torch.tensor(o._b2, dtype=f32), torch.tensor(o._eps, dtype=f32)
        one = torch.tensor(1.0, dtype=f32)
        alpha = lr * torch.sqrt(one - b2p) / (one - b1p)      # training_ops.cc ApplyAdam<CPUDevice>
        for v, g in zip(used, grads):
            if g is None:
                continue
            ms, vs = o._slots[id(v)]
            m, vv, var = sess._vars[id(ms)], sess._vars[id(vs)], sess._vars[id(v)]
            g = g.detach()
            m = m + (g - m) * (one - b1)
            vv = vv + (g * g - vv) * (one - b2)
            var = var - (m * alpha) / (torch.sqrt(vv) + eps)
            sess._vars[id(ms)], sess._vars[id(vs)], sess._vars[id(v)] = m, vv, var
        sess._vars[id(o._b1p)] = b1p * b1
        sess._vars[id(o._b2p)] = b2p * b2


train = types.SimpleNamespace(AdamOptimizer=AdamOptimizer)
initializers = types.SimpleNamespace(variables=variables_initializer, global_variables=global_variables_initializer)


def gradients(ys, xs, grad_ys=None, name=None):
    """tf.gradients for the tests of single ops (the reference itself only uses minimize)."""
    xs_l = xs if isinstance(xs, (list, tuple)) else [xs]

    def mk(i):
        class _G(Tensor):
            def _compute(self_, env):
                key = ('__grads__', id(ys), tuple(id(x) for x in xs_l))
                if key not in env:
                    leaves = []
                    for x in xs_l:
                        t = _to_t(_ev(x, env))
                        if not t.requires_grad:
                            raise RuntimeError('tfshim.gradients: feed %s as a torch tensor with requires_grad' % x)
                        leaves.append(t)
                    y = _ev(ys, env)
                    env[key] = torch.autograd.grad(y.sum(), leaves, allow_unused=True, retain_graph=True)
                return env[key][i]
        return _G('Gradient', [], None)
    out = [mk(i) for i in _py_range(len(xs_l))]
    return out


# ----------------------------------------------------------------------------------------------
# tf.contrib.slim (conv2d / pools / arg_scope / checkpoint assignment)
_ARG_SCOPES = [{}]
_CHECKPOINTS = {}
_MODEL_VARIABLES = []


def register_checkpoint(path, name_to_array):
    """Stand-in for the TF checkpoint file at ``path`` (slim variable names -> arrays)."""
    _CHECKPOINTS[path] = dict(name_to_array)


def _scoped(fn):
    def wrapper(*args, **kwargs):
        merged = dict(_ARG_SCOPES[-1].get(wrapper, {}))
        merged.update(kwargs)
        return fn(*args, **merged)
    wrapper.__name__ = fn.__name__
    wrapper._slim_fn = fn
    return wrapper


@contextlib.contextmanager
def _arg_scope(list_ops_or_scope, **kwargs):
    if isinstance(list_ops_or_scope, dict):
        new = dict(list_ops_or_scope)
    else:
        new = {k: dict(v) for k, v in _ARG_SCOPES[-1].items()}
        for op in list_ops_or_scope:
            d = dict(new.get(op, {}))
            d.update(kwargs)
            new[op] = d
    _ARG_SCOPES.append(new)
    try:
        yield new
    finally:
        _ARG_SCOPES.pop()


@_scoped
def _slim_conv2d(inputs, num_outputs, kernel_size, stride=1, padding='SAME', activation_fn=None,
                 biases_initializer='zeros', scope=None, **kw):
    kh, kw_ = (kernel_size, kernel_size) if isinstance(kernel_size, int) else kernel_size
    cin = _sshape(inputs)[-1]
    with variable_scope(scope, 'Conv'):
        w = Variable(np.zeros([kh, kw_, int(cin), num_outputs], np.float32), name='weights')
        b = Variable(np.zeros([num_outputs], np.float32), name='biases') if biases_initializer is not None else None
    _MODEL_VARIABLES.extend([w] + ([b] if b is not None else []))
    # model weights are frozen constants for the optimiser (var_list never contains them)
    y = nn.conv2d(inputs, stop_gradient(w), [1, stride, stride, 1], padding)
    if b is not None:
        y = y + stop_gradient(b)                              # nn.bias_add
    if activation_fn is not None:
        y = activation_fn(y)
    return y


@_scoped
def _slim_avg_pool2d(inputs, kernel_size, stride=2, padding='VALID', scope=None, **kw):
    k = [kernel_size] * 2 if isinstance(kernel_size, int) else list(kernel_size)
    return nn.avg_pool(inputs, [1] + k + [1], [1, stride, stride, 1], padding)


@_scoped
def _slim_max_pool2d(inputs, kernel_size, stride=2, padding='VALID', scope=None, **kw):
    k = [kernel_size] * 2 if isinstance(kernel_size, int) else list(kernel_size)
    return nn.max_pool(inputs, [1] + k + [1], [1, stride, stride, 1], padding)


def _get_model_variables(scope=None):
    return [v for v in _MODEL_VARIABLES if scope is None or v.name.startswith(scope)]


def _assign_from_checkpoint_fn(model_path, var_list, **kw):
    def init(sess):
        ck = _CHECKPOINTS[model_path]
        for v in var_list:
            sess._vars[id(v)] = torch.as_tensor(np.asarray(ck[v.name], np.float32)).clone()
    return init


contrib = types.SimpleNamespace(slim=types.SimpleNamespace(
    arg_scope=_arg_scope, conv2d=_slim_conv2d, avg_pool2d=_slim_avg_pool2d, max_pool2d=_slim_max_pool2d,
    get_model_variables=_get_model_variables, assign_from_checkpoint_fn=_assign_from_checkpoint_fn))

layers = types.SimpleNamespace(flatten=lambda x: Tensor('Flatten', [x], lambda a: a.reshape(a.shape[0], -1)))

# ----------------------------------------------------------------------------------------------
# namespaces the reference spells out
math = types.SimpleNamespace(squared_difference=squared_difference, ceil=ceil, floor=floor, log=log, exp=exp,
                             sqrt=sqrt, abs=abs, maximum=maximum, minimum=minimum, reduce_sum=reduce_sum,
                             reduce_mean=reduce_mean, reduce_max=reduce_max)
io = types.SimpleNamespace(gfile=types.SimpleNamespace(GFile=open))
gfile = io.gfile
_v1 = types.SimpleNamespace(
    placeholder=placeholder, where=where, variable_scope=variable_scope, variables_initializer=variables_initializer,
    train=train, initializers=initializers, InteractiveSession=InteractiveSession, Session=Session,
    GraphDef=GraphDef, Graph=Graph, image=image, global_variables_initializer=global_variables_initializer)
compat = types.SimpleNamespace(v1=_v1)
