"""Minimal pure-Python stand-in for the ``partio`` Python module -- the particle file I/O on either side
of the stylisation path.

The reference drivers read and write Houdini classic binary geometry (``.bgeo``) through Disney's partio
bindings (``test_smokegun.py:11-12,30-52``, ``test_chocolate.py:28-60,108-124``,
``test_smokegun_resim.py:296-320``); that C++ library is a third-party dependency absent from
``/root/reference`` and from this image.  This module restates the slice of its API the drivers call --
``create / read / write``, ``ParticlesData.addAttribute / addParticle / set / get / attributeInfo /
numParticles / numAttributes`` and the ``INT / FLOAT / VECTOR`` type tags -- over the published BGEO v5
layout (big-endian):

    int32 'Bgeo', char 'V', int32 version=5,
    int32 nPoints, nPrims, nPointGroups, nPrimGroups, nPointAttrib, nVertexAttrib, nPrimAttrib, nAttrib
    per point attribute: int16 len + name, int16 size, int32 type (0 float, 1 int, 5 vector; 4 = indexed
        string: int32 count + strings), size x int32 defaults
    per point: float32 x,y,z,w then the attribute values (4 bytes each) in declaration order
    [one particle-system primitive, "generator" primitive attribute, 0x00 0xff trailer]

Files are gzip-compressed on write (partio's default) and sniffed on read.  Host-side code only; bulk
access goes through NumPy (``ParticlesData.array``), the per-particle ``get``/``set`` calls exist for the
drivers' loops.
"""
import gzip
import struct

import numpy as np

NONE, VECTOR, FLOAT, INT, INDEXEDSTR = 0, 1, 2, 3, 4      # partio's ParticleAttributeType values
_MAGIC = ((((ord('B') << 8) | ord('g')) << 8 | ord('e')) << 8) | ord('o')
_HTYPE = {FLOAT: 0, INT: 1, VECTOR: 5}


class ParticleAttribute(object):
    def __init__(self, name, type, count, index):
        self.name, self.type, self.count, self.attributeIndex = name, type, count, index

    def __repr__(self):
        return 'ParticleAttribute(%r, type=%d, count=%d)' % (self.name, self.type, self.count)


class ParticlesData(object):
    """Struct-of-arrays particle set (one [N,count] array per attribute)."""

    def __init__(self):
        self._attrs = []
        self._data = {}
        self._n = 0

    # ---- partio API ---------------------------------------------------------------------------------
    def numParticles(self):
        return self._n

    def numAttributes(self):
        return len(self._attrs)

    def attributeInfo(self, name_or_index):
        if isinstance(name_or_index, int):
            return self._attrs[name_or_index]
        for a in self._attrs:
            if a.name == name_or_index:
                return a
        return None

    def addAttribute(self, name, type, count):
        if self.attributeInfo(name) is not None:
            raise ValueError('attribute %r exists' % name)
        a = ParticleAttribute(name, type, int(count), len(self._attrs))
        self._attrs.append(a)
        self._data[name] = np.zeros([max(self._n, 16), a.count], np.int32 if type == INT else np.float32)
        return a

    def addParticle(self):
        return self.addParticles(1)

    def addParticles(self, count):
        first = self._n
        self._n += int(count)
        for k, arr in self._data.items():
            if arr.shape[0] < self._n:
                grown = np.zeros([max(self._n, 2 * arr.shape[0]), arr.shape[1]], arr.dtype)
                grown[:first] = arr[:first]
                self._data[k] = grown
        return first

    def set(self, attr, index, values):
        self._data[attr.name][index] = values

    def get(self, attr, index):
        row = self._data[attr.name][index]
        return tuple(int(v) for v in row) if attr.type == INT else tuple(float(v) for v in row)

    # ---- bulk access (not in partio's Python API; used by lnst.drivers) -------------------------------
    def array(self, name):
        return self._data[name][:self._n]

    def setArray(self, name, values):
        values = np.asarray(values)
        if values.ndim == 1:
            values = values[:, None]
        if values.shape[0] > self._n:
            self.addParticles(values.shape[0] - self._n)
        self._data[name][:values.shape[0]] = values


def create():
    return ParticlesData()


def _hstr(s):
    b = s.encode('ascii')
    return struct.pack('>h', len(b)) + b


def write(path, pt, compressed=True):
    """partio.write for ``.bgeo`` (BGEO.cpp writeBGEO): 'position' becomes the point coordinates (w = 1)."""
    if not str(path).endswith('.bgeo'):
        raise ValueError('only .bgeo is supported: %s' % path)
    n = pt.numParticles()
    attrs = [a for a in pt._attrs if a.name != 'position']
    for a in attrs:
        if a.type not in _HTYPE:
            raise ValueError('attribute %r: unsupported type %d' % (a.name, a.type))
    out = [struct.pack('>icii', _MAGIC, b'V', 5, n), struct.pack('>iiiiiii', 1, 0, 0, len(attrs), 0, 1, 0)]
    for a in attrs:
        out.append(_hstr(a.name) + struct.pack('>Hi', a.count, _HTYPE[a.type]) + struct.pack('>%di' % a.count, *([0] * a.count)))
    width = 4 + sum(a.count for a in attrs)
    rec = np.zeros([n, width], '>u4')
    pos = pt.array('position') if pt.attributeInfo('position') is not None else np.zeros([n, 3], np.float32)
    k = min(pos.shape[1], 3)                               # a 2-component position (test_dambreak2d.py:101) gets z = 0
    rec[:, :k] = pos[:, :k].astype('>f4').view('>u4')
    rec[:, 3] = np.ones(n, '>f4').view('>u4')
    col = 4
    for a in attrs:
        dt = '>i4' if a.type == INT else '>f4'
        rec[:, col:col + a.count] = pt.array(a.name).astype(dt).view('>u4')
        col += a.count
    out.append(rec.tobytes())
    out.append(_hstr('generator') + struct.pack('>hii', 1, 4, 1) + _hstr('papi'))
    out.append(struct.pack('>ii', 0x8000, n))
    out.append(np.arange(n).astype('>i4' if n > (1 << 16) else '>u2').tobytes())
    out.append(struct.pack('>i', 0) + b'\x00\xff')
    blob = b''.join(out)
    with (gzip.open(path, 'wb') if compressed else open(path, 'wb')) as f:
        f.write(blob)


def read(path):
    """partio.read for ``.bgeo`` (BGEO.cpp readBGEO): point attributes only, primitives are ignored."""
    with open(path, 'rb') as f:
        blob = f.read()
    if blob[:2] == b'\x1f\x8b':
        blob = gzip.decompress(blob)
    magic, vchar, version, n = struct.unpack_from('>icii', blob, 0)
    if magic != _MAGIC or vchar != b'V' or version != 5:
        raise ValueError('%s: not a BGEO v5 file' % path)
    off = struct.calcsize('>icii')
    _nprims, _npg, _nprg, n_pattr, _nv, _npa, _na = struct.unpack_from('>iiiiiii', blob, off)
    off += 28
    pt = ParticlesData()
    pt.addAttribute('position', VECTOR, 3)
    layout = []
    for _ in range(n_pattr):
        ln, = struct.unpack_from('>h', blob, off)
        name = blob[off + 2:off + 2 + ln].decode('ascii')
        off += 2 + ln
        size, htype = struct.unpack_from('>Hi', blob, off)
        off += 6
        if htype == 4:                                         # indexed strings: table is skipped, indices kept as INT
            cnt, = struct.unpack_from('>i', blob, off)
            off += 4
            for _s in range(cnt):
                sl, = struct.unpack_from('>h', blob, off)
                off += 2 + sl
            size, typ = 1, INT
        elif htype in (0, 1, 5):
            off += 4 * size
            typ = {0: FLOAT, 1: INT, 5: VECTOR}[htype]
        else:
            raise ValueError('%s: unsupported Houdini attribute type %d (%s)' % (path, htype, name))
        layout.append((pt.addAttribute(name, typ, size), size))
    width = 4 + sum(s for _, s in layout)
    rec = np.frombuffer(blob, '>u4', count=n * width, offset=off).reshape(n, width)
    pt.addParticles(n)
    pt.setArray('position', rec[:, :3].view('>f4').astype(np.float32))
    col = 4
    for a, size in layout:
        raw = rec[:, col:col + size]
        pt.setArray(a.name, raw.view('>i4').astype(np.int32) if a.type == INT else raw.view('>f4').astype(np.float32))
        col += size
    return pt
