"""ASCII Open Inventor (.iv) triangle loader with the semantics of the reference's ivcon path.

Reference: src/libopenrave-core/xmlreaders-core.cpp:469-478 (CreateTriMeshFromFile falls back to
`ivcon::ReadFile`) and 3rdparty/ivcon/ivcon.cpp:2658-2716 (ReadFile), :8822-8869 (Coordinate3.point
parsed with sscanf("%f") into FLOAT storage).  What that path does, and what this module restates:

  * every `Coordinate3 { point [...] }` list in the file is appended, in file order, to one vertex
    array of single-precision floats; no node transform other than MatrixTransform is applied
    (Translation/Rotation/Cylinder nodes such as the ones in models/barrett/palm.iv are ignored);
  * if the file has no IndexedFaceSet faces, every three consecutive vertices form a triangle
    (ReadFile, `face_num == 0` branch); otherwise faces are fan-triangulated.

Binary Inventor 2.1 files (models/furniture/computer.iv in data/wamtest1.env.xml) cannot be read by ivcon; the reference
reads them through Coin (`ivmodelloader`, xmlreaders-core.cpp:451-467), whose triangle callback walks every shape.  The
only shapes in the reference's binary models are `IndexedTriangleStripSet`s carrying their own `VertexProperty` and no
transform nodes, so `_load_iv_binary` restates just that: big-endian MFVec3f `vertex` + MFInt32 `coordIndex` per shape,
strips (i, i+1, i+2) cut at -1, degenerate triangles kept (verdict-neutral).  No Coin here => this mesh is "parity
unpinned" at the vertex level; it is pinned only through the scene-level checks in tests/.

Load-time tooling (scene ingestion is not part of the per-config hot path).
"""
import ctypes
import re

import numpy as np

_libc = ctypes.CDLL(None)
_libc.strtof.restype = ctypes.c_float
_libc.strtof.argtypes = [ctypes.c_char_p, ctypes.c_void_p]

_POINT_RE = re.compile(r"Coordinate3\s*\{\s*point\s*\[([^\]]*)\]", re.S)
_INDEX_RE = re.compile(r"IndexedFaceSet\s*\{[^}]*?coordIndex\s*\[([^\]]*)\]", re.S)


def _parse_floats32(text):
    toks = text.replace(",", " ").split()
    out = np.empty(len(toks), dtype=np.float32)
    strtof = _libc.strtof
    for i, t in enumerate(toks):
        out[i] = strtof(t.encode(), None)   # single rounding, like sscanf("%f")
    return out


_BIN_VERTEX = re.compile(rb"\x00\x00\x00\x06vertex\x00\x00")
_BIN_INDEX = re.compile(rb"\x00\x00\x00\x0acoordIndex\x00\x00")


def _load_iv_binary(path, data):
    verts, tris, base = [], [], 0
    vpos = [m.end() for m in _BIN_VERTEX.finditer(data)]
    ipos = [m.end() for m in _BIN_INDEX.finditer(data)]
    if not vpos or len(vpos) != len(ipos):
        raise ValueError("%s: binary Inventor file without VertexProperty/coordIndex shape pairs" % path)
    for vo, io in zip(vpos, ipos):
        nv = int(np.frombuffer(data, dtype=">i4", count=1, offset=vo)[0])
        v = np.frombuffer(data, dtype=">f4", count=3 * nv, offset=vo + 4).astype(np.float32).reshape(-1, 3)
        ni = int(np.frombuffer(data, dtype=">i4", count=1, offset=io)[0])
        idx = np.frombuffer(data, dtype=">i4", count=ni, offset=io + 4)
        strip = []
        for k in list(idx) + [-1]:
            if k < 0:
                for j in range(len(strip) - 2):
                    a, b, c = strip[j], strip[j + 1], strip[j + 2]
                    tris.append((base + a, base + c, base + b) if (j & 1) else (base + a, base + b, base + c))
                strip = []
            else:
                if k >= nv:
                    raise ValueError("%s: coordIndex out of range" % path)
                strip.append(int(k))
        verts.append(v)
        base += nv
    return np.concatenate(verts), np.asarray(tris, dtype=np.int32).reshape(-1, 3)


def load_iv(path):
    """Return (vertices float32 [nv,3], indices int32 [nt,3])."""
    with open(path, "rb") as f:
        head = f.read(64)
        if head.startswith(b"#Inventor") and b"binary" in head.split(b"\n")[0]:
            return _load_iv_binary(path, head + f.read())
    with open(path, "r", errors="replace") as f:
        text = f.read()
    if not text.startswith("#Inventor"):
        raise ValueError("%s: not an ASCII Inventor file" % path)
    chunks = [_parse_floats32(m.group(1)) for m in _POINT_RE.finditer(text)]
    if not chunks:
        raise ValueError("%s: no Coordinate3 points" % path)
    # ivcon carries a partially filled xyz triple across lists only within one list; every list in
    # the reference's model files holds whole triples, so a remainder is a malformed file.
    for c in chunks:
        if c.size % 3:
            raise ValueError("%s: Coordinate3 list is not a multiple of 3 floats" % path)
    verts = np.concatenate(chunks).reshape(-1, 3)

    faces = []
    for m in _INDEX_RE.finditer(text):
        cur = []
        for tok in m.group(1).replace(",", " ").split():
            v = int(tok)
            if v < 0:
                if len(cur) >= 3:
                    faces.append(cur)
                cur = []
            else:
                cur.append(v)
        if len(cur) >= 3:
            faces.append(cur)
    if not faces:
        n = verts.shape[0] // 3
        idx = np.arange(n * 3, dtype=np.int32).reshape(n, 3)
    else:
        tri = []
        for face in faces:
            for k in range(2, len(face)):
                tri.append((face[0], face[k - 1], face[k]))
        idx = np.asarray(tri, dtype=np.int32)
    return verts, idx
