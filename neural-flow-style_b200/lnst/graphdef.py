"""TensorFlow ``GraphDef`` files without TensorFlow: a reader (and a writer, for synthetic test graphs) of the
protobuf wire format, restricted to what a frozen inference graph such as ``tensorflow_inception_graph.pb``
(inception5h; reference ``styler_base.py:19-31``: ``GraphDef.ParseFromString`` + ``tf.import_graph_def``) holds:

    GraphDef   { repeated NodeDef node = 1; }
    NodeDef    { string name = 1; string op = 2; repeated string input = 3; map<string, AttrValue> attr = 5; }
    AttrValue  { ListValue list = 1; bytes s = 2; int64 i = 3; float f = 4; bool b = 5; DataType type = 6;
                 TensorShapeProto shape = 7; TensorProto tensor = 8; }
    ListValue  { repeated bytes s = 2; repeated int64 i = 3; repeated float f = 4; ... }
    TensorProto{ DataType dtype = 1; TensorShapeProto tensor_shape = 2; bytes tensor_content = 4;
                 repeated float float_val = 5; repeated int32 int_val = 7; }
    TensorShapeProto { repeated Dim dim = 2 { int64 size = 1; } }

Host-side parsing only; the tensors end up as NumPy arrays.
"""
import struct

import numpy as np

DT_FLOAT, DT_INT32 = 1, 3


# ---- wire format -------------------------------------------------------------------------------------
def _varint(buf, pos):
    out = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7f) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def _fields(buf):
    """Yield (field number, wire type, value) for one message; length-delimited values are memoryviews."""
    pos, end = 0, len(buf)
    while pos < end:
        key, pos = _varint(buf, pos)
        field, wt = key >> 3, key & 7
        if wt == 0:
            val, pos = _varint(buf, pos)
        elif wt == 1:
            val, pos = bytes(buf[pos:pos + 8]), pos + 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            val, pos = buf[pos:pos + ln], pos + ln
        elif wt == 5:
            val, pos = bytes(buf[pos:pos + 4]), pos + 4
        else:
            raise ValueError('unsupported protobuf wire type %d' % wt)
        yield field, wt, val


def _signed(v):
    return v - (1 << 64) if v >= (1 << 63) else v


def _packed_varints(wt, val):
    if wt == 0:
        return [_signed(val)]
    out, pos = [], 0
    while pos < len(val):
        v, pos = _varint(val, pos)
        out.append(_signed(v))
    return out


def _packed_floats(wt, val):
    if wt == 5:
        return [struct.unpack('<f', val)[0]]
    return list(np.frombuffer(bytes(val), '<f4'))


def _shape(buf):
    dims = []
    for f, wt, v in _fields(buf):
        if f == 2:
            size = 0
            for f2, _, v2 in _fields(v):
                if f2 == 1:
                    size = _signed(v2)
            dims.append(size)
    return dims


def _tensor(buf):
    dtype, shape, content, fvals, ivals = DT_FLOAT, [], None, [], []
    for f, wt, v in _fields(buf):
        if f == 1:
            dtype = v
        elif f == 2:
            shape = _shape(v)
        elif f == 4:
            content = bytes(v)
        elif f == 5:
            fvals += _packed_floats(wt, v)
        elif f == 7:
            ivals += _packed_varints(wt, v)
    np_dt = {DT_FLOAT: '<f4', DT_INT32: '<i4'}.get(dtype)
    if np_dt is None:
        return None                                            # dtypes the loss networks never read
    n = int(np.prod(shape)) if shape else 1
    if content is not None:
        arr = np.frombuffer(content, np_dt)
    else:
        vals = fvals if dtype == DT_FLOAT else ivals
        arr = np.asarray(vals, np_dt)
        if arr.size == 1 and n > 1:                            # a single value stands for a constant-filled tensor
            arr = np.full(n, arr[0], np_dt)
        elif arr.size == 0:
            arr = np.zeros(n, np_dt)
    return arr.reshape(shape).copy()


def _attr(buf):
    for f, wt, v in _fields(buf):
        if f == 2:
            return bytes(v)
        if f == 3:
            return _signed(v)
        if f == 4:
            return struct.unpack('<f', v)[0]
        if f == 5:
            return bool(v)
        if f == 6:
            return ('type', v)
        if f == 7:
            return _shape(v)
        if f == 8:
            return _tensor(v)
        if f == 1:
            s, i, fl = [], [], []
            for f2, wt2, v2 in _fields(v):
                if f2 == 2:
                    s.append(bytes(v2))
                elif f2 == 3:
                    i += _packed_varints(wt2, v2)
                elif f2 == 4:
                    fl += _packed_floats(wt2, v2)
            return s or i or fl
    return None


class Node(object):
    __slots__ = ('name', 'op', 'inputs', 'attr')

    def __init__(self, name, op, inputs=(), attr=None):
        self.name, self.op, self.inputs, self.attr = name, op, list(inputs), dict(attr or {})

    def __repr__(self):
        return 'Node(%s %s <- %s)' % (self.op, self.name, self.inputs)


def parse(blob):
    """GraphDef bytes -> list of ``Node`` in file order."""
    nodes = []
    for f, _, v in _fields(memoryview(blob)):
        if f != 1:
            continue
        name = op = ''
        inputs, attr = [], {}
        for f2, _, v2 in _fields(v):
            if f2 == 1:
                name = bytes(v2).decode()
            elif f2 == 2:
                op = bytes(v2).decode()
            elif f2 == 3:
                inputs.append(bytes(v2).decode())
            elif f2 == 5:
                key, val = None, None
                for f3, _, v3 in _fields(v2):
                    if f3 == 1:
                        key = bytes(v3).decode()
                    elif f3 == 2:
                        val = _attr(v3)
                attr[key] = val
        nodes.append(Node(name, op, inputs, attr))
    return nodes


def load(path):
    with open(path, 'rb') as f:
        return parse(f.read())


# ---- writer (synthetic graphs for tests / bench; same subset) -------------------------------------------
def _enc_varint(v):
    v &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = v & 0x7f
        v >>= 7
        out.append(b | (0x80 if v else 0))
        if not v:
            return bytes(out)


def _ld(field, payload):
    return _enc_varint(field << 3 | 2) + _enc_varint(len(payload)) + payload


def _enc_shape(dims):
    return b''.join(_ld(2, _enc_varint(1 << 3) + _enc_varint(int(d))) for d in dims)


def _enc_attr(v):
    if isinstance(v, np.ndarray):
        dt = DT_INT32 if v.dtype.kind in 'iu' else DT_FLOAT
        body = (_enc_varint(1 << 3) + _enc_varint(dt) + _ld(2, _enc_shape(v.shape)) +
                _ld(4, np.ascontiguousarray(v, '<i4' if dt == DT_INT32 else '<f4').tobytes()))
        return _ld(8, body)
    if isinstance(v, bool):
        return _enc_varint(5 << 3) + _enc_varint(int(v))
    if isinstance(v, int):
        return _enc_varint(3 << 3) + _enc_varint(v)
    if isinstance(v, float):
        return _enc_varint(4 << 3 | 5) + struct.pack('<f', v)
    if isinstance(v, (bytes, str)):
        return _ld(2, v.encode() if isinstance(v, str) else v)
    if isinstance(v, (list, tuple)):                           # list(i), packed
        return _ld(1, _ld(3, b''.join(_enc_varint(int(i)) for i in v)))
    raise TypeError('cannot encode attribute %r' % (v,))


def serialize(nodes):
    """list of ``Node`` -> GraphDef bytes."""
    out = []
    for n in nodes:
        body = _ld(1, n.name.encode()) + _ld(2, n.op.encode()) + b''.join(_ld(3, i.encode()) for i in n.inputs)
        for k in sorted(n.attr):
            body += _ld(5, _ld(1, k.encode()) + _ld(2, _enc_attr(n.attr[k])))
        out.append(_ld(1, body))
    return b''.join(out)
