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

ASCII files that pull other files in through `File { name "..." }` nodes (models/objects/ketchup.iv in
data/wamtest1.env.xml) carry nothing ivcon can read; Coin resolves the include relative to the including file and applies
the transform nodes accumulated inside the enclosing `Separator` (RotationXYZ, Transform{translation, rotation,
scaleFactor}, Translation, Rotation, Scale; each right-multiplies the current matrix, SoSeparator restores it on exit).
`_load_iv_graph` restates that for exactly those node types ("parity unpinned" at the vertex level, as above).

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


_TOKEN_RE = re.compile(r'"[^"]*"|[{}\[\]]|[^\s{}\[\]]+')


def _axis_angle(axis, angle):
    axis = np.asarray(axis, dtype=np.float64)
    n = np.linalg.norm(axis)
    if n == 0:
        return np.eye(3)
    x, y, z = axis / n
    c, s = np.cos(angle), np.sin(angle)
    C = 1 - c
    return np.array([[c + x * x * C, x * y * C - z * s, x * z * C + y * s],
                     [y * x * C + z * s, c + y * y * C, y * z * C - x * s],
                     [z * x * C - y * s, z * y * C + x * s, c + z * z * C]])


def _load_iv_graph(path, text):
    """Separator / transform / File walker (Coin semantics) for ASCII files that include other files."""
    import os
    toks = _TOKEN_RE.findall(re.sub(r"#[^\n]*", "", text))
    verts, tris, base = [], [], 0
    stack, cur = [], np.eye(4)
    i = 0

    def fields(j):
        """tokens of the node body starting after '{' at j -> (dict field -> [tokens], index after '}')."""
        out, depth, name = {}, 1, None
        while j < len(toks) and depth:
            t = toks[j]
            if t == "{":
                depth += 1
            elif t == "}":
                depth -= 1
            elif depth == 1:
                if re.match(r"^[A-Za-z_]\w*$", t) and (name is None or t[0].islower() and not out[name] or
                                                      t in ("translation", "rotation", "scaleFactor", "axis", "angle", "name")):
                    name = t
                    out[name] = []
                elif name is not None:
                    out[name].append(t)
            j += 1
        return out, j

    while i < len(toks):
        t = toks[i]
        if t == "Separator" and i + 1 < len(toks) and toks[i + 1] == "{":
            stack.append(cur.copy())
            i += 2
            continue
        if t == "}":
            if stack:
                cur = stack.pop()
            i += 1
            continue
        if i + 1 < len(toks) and toks[i + 1] == "{":
            f, nxt = fields(i + 2)
            M = np.eye(4)
            if t == "RotationXYZ":
                ax = {"X": (1, 0, 0), "Y": (0, 1, 0), "Z": (0, 0, 1)}[f.get("axis", ["X"])[0]]
                M[:3, :3] = _axis_angle(ax, float(f.get("angle", ["0"])[0]))
            elif t in ("Transform", "Translation", "Rotation", "Scale"):
                if "translation" in f:
                    M[:3, 3] = [float(x) for x in f["translation"][:3]]
                if "rotation" in f:
                    r = [float(x) for x in f["rotation"][:4]]
                    M[:3, :3] = _axis_angle(r[:3], r[3])
                if "scaleFactor" in f:
                    M[:3, :3] = M[:3, :3] @ np.diag([float(x) for x in f["scaleFactor"][:3]])
            elif t == "File":
                sub = os.path.join(os.path.dirname(path), f["name"][0].strip('"'))
                v, idx = load_iv(sub)
                v = (v.astype(np.float64) @ cur[:3, :3].T + cur[:3, 3]).astype(np.float32)
                verts.append(v)
                tris.append(idx + base)
                base += len(v)
            cur = cur @ M
            i = nxt
            continue
        i += 1
    if not verts:
        raise ValueError("%s: no Coordinate3 points" % path)
    return np.concatenate(verts), np.concatenate(tris).astype(np.int32)


def load_iv(path):
    """Return (vertices float32 [nv,3], indices int32 [nt,3])."""
    with open(path, "rb") as f:
        head = f.read(64)
        if head.startswith(b"#Inventor") and b"binary" in head.split(b"\n")[0]:
            return _load_iv_binary(path, head + f.read())
    with open(path, "r", errors="replace") as f:
        text = f.read()
    if not (text.startswith("#Inventor") or text.startswith("#VRML V1.0 ascii")):     # VRML 1.0 is Inventor's ASCII grammar
        raise ValueError("%s: not an ASCII Inventor file" % path)
    chunks = [_parse_floats32(m.group(1)) for m in _POINT_RE.finditer(text)]
    if not chunks and re.search(r"\bFile\s*\{", text):
        return _load_iv_graph(path, text)
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
