"""TensorFlow V1 checkpoint files (``vgg_19.ckpt`` of the slim model zoo, reference ``README.md:25`` /
``vgg.py:115-120``: ``slim.assign_from_checkpoint_fn(model_path, slim.get_model_variables('vgg_19'))``) without
TensorFlow.

A V1 checkpoint is one file written by TF's ``TensorSliceWriter``: a LevelDB-style sorted string table

    [data block]* [metaindex block] [index block] [footer: 2 block handles, padding to 40 bytes, magic 0xdb4775248b80fb57]
    block   = entries, restart array (uint32 each), uint32 restart count; followed by 1 type byte (0 raw, 1 snappy)
              and a masked crc32c
    entry   = varint shared, varint non_shared, varint value_len, key suffix, value   (keys are prefix-compressed)

whose values are ``SavedTensorSlices`` protos: the entry with the empty key carries ``meta`` (name, shape, dtype of
every tensor), every other entry carries one slice of one tensor as a ``TensorProto`` with packed ``float_val``.
``read`` returns ``{name: float32 ndarray}``; partitioned tensors are reassembled from their slices.  ``write``
produces the same layout (uncompressed, valid CRCs) -- used by the tests, and handy to hand seeded weights to a
TensorFlow installation.

PARITY NOTE: restated from the published table / proto definitions; no TensorFlow-written file exists in this
environment to read back, so the reader is checked against this writer and against the protobuf library only.
"""
import struct

import numpy as np

from .graphdef import _fields, _varint, _packed_floats, _packed_varints, _shape, _enc_varint, _ld, _enc_shape

MAGIC = 0xdb4775248b80fb57
DT_FLOAT, DT_INT32 = 1, 3


# ---- crc32c (Castagnoli), masked as LevelDB does -------------------------------------------------------
def _crc_table():
    tab = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        tab.append(c)
    return tab


_CRC = _crc_table()


def crc32c(data, crc=0):
    crc ^= 0xFFFFFFFF
    for b in data:
        crc = _CRC[(crc ^ b) & 0xFF] ^ (crc >> 8)
    return crc ^ 0xFFFFFFFF


def _mask(crc):
    return (((crc >> 15) | (crc << 17)) + 0xa282ead8) & 0xFFFFFFFF


# ---- snappy (raw format) decoder: TF writes uncompressed tables, other writers may not ------------------
def _snappy(buf):
    n, pos = _varint(buf, 0)
    out = bytearray()
    while pos < len(buf):
        tag = buf[pos]
        pos += 1
        kind = tag & 3
        if kind == 0:
            ln = tag >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(bytes(buf[pos:pos + nb]), 'little')
                pos += nb
            ln += 1
            out += buf[pos:pos + ln]
            pos += ln
            continue
        if kind == 1:
            ln, off = ((tag >> 2) & 7) + 4, ((tag >> 5) << 8) | buf[pos]
            pos += 1
        elif kind == 2:
            ln, off = (tag >> 2) + 1, int.from_bytes(bytes(buf[pos:pos + 2]), 'little')
            pos += 2
        else:
            ln, off = (tag >> 2) + 1, int.from_bytes(bytes(buf[pos:pos + 4]), 'little')
            pos += 4
        if off == 0 or off > len(out):
            raise ValueError('corrupt snappy block')
        for _ in range(ln):                                    # overlapping copies are legal
            out.append(out[-off])
    if len(out) != n:
        raise ValueError('corrupt snappy block (length)')
    return memoryview(bytes(out))


# ---- table reader -------------------------------------------------------------------------------------------
def _handle(buf, pos):
    off, pos = _varint(buf, pos)
    size, pos = _varint(buf, pos)
    return off, size, pos


def _block(blob, off, size):
    raw = blob[off:off + size]
    kind = blob[off + size]
    if kind == 1:
        raw = _snappy(raw)
    elif kind != 0:
        raise ValueError('unknown block compression %d' % kind)
    n_restarts, = struct.unpack_from('<I', raw, len(raw) - 4)
    end = len(raw) - 4 - 4 * n_restarts
    pos, key = 0, b''
    while pos < end:
        shared, pos = _varint(raw, pos)
        non_shared, pos = _varint(raw, pos)
        vlen, pos = _varint(raw, pos)
        key = key[:shared] + bytes(raw[pos:pos + non_shared])
        pos += non_shared
        yield key, raw[pos:pos + vlen]
        pos += vlen


def _entries(blob):
    if len(blob) < 48 or struct.unpack_from('<Q', blob, len(blob) - 8)[0] != MAGIC:
        raise ValueError('not a TensorFlow V1 checkpoint (table magic missing)')
    foot = blob[len(blob) - 48:]
    _mo, _ms, pos = _handle(foot, 0)
    io, isz, _ = _handle(foot, pos)
    for _key, h in _block(blob, io, isz):
        off, size, _ = _handle(h, 0)
        for kv in _block(blob, off, size):
            yield kv


def _slice_extents(buf):
    ext = []
    for f, _, v in _fields(buf):
        if f == 1:
            start, length = 0, None
            for f2, wt2, v2 in _fields(v):
                if f2 == 1:
                    start = _packed_varints(wt2, v2)[0]
                elif f2 == 2:
                    length = _packed_varints(wt2, v2)[0]
            ext.append((start, length))
    return ext


def read(path, names=None):
    """{tensor name: ndarray} of a V1 checkpoint file.  ``names``: only these (a predicate or a container)."""
    with open(path, 'rb') as f:
        blob = memoryview(f.read())
    want = (lambda n: True) if names is None else (names if callable(names) else (lambda n: n in names))
    meta, out = {}, {}
    for key, val in _entries(blob):
        for f, _, v in _fields(val):
            if f == 1:                                         # SavedTensorSliceMeta
                for f2, _, v2 in _fields(v):
                    if f2 != 1:
                        continue
                    name, shape, dtype = '', [], DT_FLOAT
                    for f3, _, v3 in _fields(v2):
                        if f3 == 1:
                            name = bytes(v3).decode()
                        elif f3 == 2:
                            shape = _shape(v3)
                        elif f3 == 3:
                            dtype = v3
                    meta[name] = (shape, dtype)
            elif f == 2:                                       # SavedSlice
                name, ext, data = '', [], None
                for f2, _, v2 in _fields(v):
                    if f2 == 1:
                        name = bytes(v2).decode()
                    elif f2 == 2:
                        ext = _slice_extents(v2)
                    elif f2 == 3:
                        data = v2
                if not want(name) or data is None:
                    continue
                shape, dtype = meta.get(name, (None, DT_FLOAT))
                vals = []
                for f3, wt3, v3 in _fields(data):
                    if f3 == 5 and dtype == DT_FLOAT:
                        vals.append(np.frombuffer(bytes(v3), '<f4') if wt3 == 2 else np.asarray(_packed_floats(wt3, v3), '<f4'))
                    elif f3 == 7 and dtype == DT_INT32:
                        vals.append(np.asarray(_packed_varints(wt3, v3), '<i4'))
                    elif f3 == 4:
                        vals.append(np.frombuffer(bytes(v3), '<f4' if dtype == DT_FLOAT else '<i4'))
                if not vals:
                    continue
                arr = np.concatenate(vals) if len(vals) > 1 else vals[0]
                if shape is None:
                    out[name] = arr
                    continue
                full = out.get(name)
                if full is None:
                    full = out[name] = np.zeros(shape, arr.dtype)
                idx, sub = [], []
                for d, size in enumerate(shape):
                    start, length = ext[d] if d < len(ext) else (0, None)
                    length = size - start if length is None else length
                    idx.append(slice(start, start + length))
                    sub.append(length)
                full[tuple(idx)] = arr.reshape(sub)
    return out


# ---- table writer (uncompressed, one data block per tensor) ----------------------------------------------------
def _enc_block(entries):
    body = bytearray()
    for key, val in entries:                                   # every entry is its own restart point (no sharing)
        body += _enc_varint(0) + _enc_varint(len(key)) + _enc_varint(len(val)) + key + val
    offs, pos = [], 0
    for key, val in entries:
        offs.append(pos)
        pos += len(_enc_varint(0) + _enc_varint(len(key)) + _enc_varint(len(val))) + len(key) + len(val)
    if not offs:
        offs = [0]
    body += b''.join(struct.pack('<I', o) for o in offs) + struct.pack('<I', len(offs))
    return bytes(body)


def _ordered_key(name):
    """EncodeTensorNameSlice for a full slice of a rank-0 view: only ordering matters to readers; names sort."""
    return b'\x00' + name.encode().replace(b'\x00', b'\x00\xff') + b'\x00\x01'


def write(path, tensors):
    """{name: float32/int32 ndarray} -> V1 checkpoint file."""
    items = sorted((k, np.asarray(v)) for k, v in tensors.items())
    meta = b''
    for name, a in items:
        dt = DT_INT32 if a.dtype.kind in 'iu' else DT_FLOAT
        full = b''.join(_ld(1, b'') for _ in a.shape)          # TensorSliceProto: one empty Extent per dim = full
        meta += _ld(1, _ld(1, name.encode()) + _ld(2, _enc_shape(a.shape)) + _enc_varint(3 << 3) + _enc_varint(dt) +
                    _ld(4, full))
    entries = [(b'', _ld(1, meta))]
    for name, a in items:
        dt = DT_INT32 if a.dtype.kind in 'iu' else DT_FLOAT
        if dt == DT_FLOAT:
            payload = _ld(5, np.ascontiguousarray(a, '<f4').tobytes())
        else:
            payload = _ld(7, b''.join(_enc_varint(int(x)) for x in a.reshape(-1)))
        tproto = _enc_varint(1 << 3) + _enc_varint(dt) + _ld(2, _enc_shape(a.shape)) + payload
        sl = b''.join(_ld(1, b'') for _ in a.shape)
        entries.append((_ordered_key(name), _ld(2, _ld(1, name.encode()) + _ld(2, sl) + _ld(3, tproto))))
    out, index = bytearray(), []

    def emit(block):
        off = len(out)
        out.extend(block)
        out.extend(b'\x00' + struct.pack('<I', _mask(crc32c(b'\x00', crc32c(block)))))
        return off, len(block)

    for key, val in entries:
        off, size = emit(_enc_block([(key, val)]))
        index.append((key + b'\x00', _enc_varint(off) + _enc_varint(size)))    # separator >= last key of the block
    mo, ms = emit(_enc_block([]))
    io, isz = emit(_enc_block(index))
    foot = _enc_varint(mo) + _enc_varint(ms) + _enc_varint(io) + _enc_varint(isz)
    out.extend(foot + b'\x00' * (40 - len(foot)) + struct.pack('<Q', MAGIC))
    with open(path, 'wb') as f:
        f.write(bytes(out))
