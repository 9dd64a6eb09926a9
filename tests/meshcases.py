"""Meshes and transforms for the mesh-ingest parity tests (read_ply's face / vertex-normal passes + Mesh's constructor,
reference base/PlyReader.cpp:487-531, shapes/Triangle.h:25-51)."""
from __future__ import annotations

import numpy as np

from simplepath_b200 import scenes


def transform(seed: int = 3) -> np.ndarray:
    """object_to_world as 12 floats c0.xyz c1.xyz c2.xyz affine.xyz: a rotation about a skew axis, non-uniform scale, offset."""
    rng = np.random.default_rng(seed)
    axis = rng.normal(size=3)
    axis /= np.linalg.norm(axis)
    a = 0.7
    k = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    rot = np.eye(3) + np.sin(a) * k + (1 - np.cos(a)) * (k @ k)
    m = rot @ np.diag([10.0, 7.5, 12.0])
    return np.concatenate([m[:, 0], m[:, 1], m[:, 2], [0.3, -1.25, 4.0]]).astype(np.float32)


def mesh(n_tris: int = 3000, seed: int = 1, fan: int = 200) -> tuple[np.ndarray, np.ndarray]:
    """A bumpy sphere plus what the reader has to cope with: zero-area faces (repeated index, collinear and coincident
    vertices), vertices no face uses, one vertex shared by a large fan, faces in shuffled order."""
    rng = np.random.default_rng(seed)
    v, f = scenes.bumpy_sphere(n_tris, (-0.1, 0.03, -0.06), (0.06, 0.19, 0.06))
    v = np.asarray(v, dtype=np.float32)
    f = np.asarray(f, dtype=np.uint32)
    nv = len(v)
    extra_v = np.array([[0, 0, 0], [1, 0, 0], [2, 0, 0], [0.5, 0.25, 0.125], [0.5, 0.25, 0.125], [9, 9, 9]], dtype=np.float32)
    v = np.concatenate([v, extra_v])
    degenerate = np.array([[nv, nv + 1, nv + 2],       # collinear
                           [nv + 3, nv + 4, 0],        # two coincident vertices
                           [5, 5, 9],                  # repeated index
                           [7, 7, 7]], dtype=np.uint32)
    hub = nv + 5                                       # a fan of 200 triangles around one vertex
    ring = rng.integers(0, nv, fan + 1).astype(np.uint32)
    fan_faces = np.stack([np.full(fan, hub, dtype=np.uint32), ring[:-1], ring[1:]], axis=1)
    fan_faces = fan_faces[(fan_faces[:, 1] != fan_faces[:, 2])]
    f = np.concatenate([f, degenerate, fan_faces])
    f = f[rng.permutation(len(f))]
    return v, f


def ulp_distance(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Distance in units of the last place between float32 arrays (same sign assumed where it matters)."""
    ia = np.ascontiguousarray(a, dtype=np.float32).view(np.int32).astype(np.int64)
    ib = np.ascontiguousarray(b, dtype=np.float32).view(np.int32).astype(np.int64)
    ia = np.where(ia < 0, -(ia & 0x7FFFFFFF), ia)
    ib = np.where(ib < 0, -(ib & 0x7FFFFFFF), ib)
    return np.abs(ia - ib)


def normal_condition(v: np.ndarray, f: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """How ill-conditioned each vertex normal is: (number of face normals summed) / |their sum| >= 1 — last-bit differences
    in the unit summands reach the normalised result amplified by this ratio.  Returns (per vertex, per corner of the faces
    with non-zero area in face order), computed in float64."""
    p = v.astype(np.float64)[f.astype(np.int64)]
    n = np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0])
    area = np.linalg.norm(n, axis=1)
    keep = area > 0
    unit = n[keep] / area[keep, None]
    fk = f[keep].astype(np.int64)
    total = np.zeros((len(v), 3))
    count = np.zeros(len(v))
    for k in range(3):
        np.add.at(total, fk[:, k], unit)
        np.add.at(count, fk[:, k], 1.0)
    length = np.linalg.norm(total, axis=1)
    cond = np.where(length > 0, count / np.maximum(length, 1e-300), 1.0)
    cond = np.maximum(cond, 1.0)
    return cond, cond[fk]


def stl_soup(n_tris: int = 3000, seed: int = 2) -> tuple[np.ndarray, np.ndarray]:
    """Triangle soup + stored normals as a binary STL holds them: (corners [t, 3, 3], normals [t, 3]).  Stored normals are a
    mix of unit normals, zero vectors (the reader falls back to the cross product), arbitrary non-unit vectors (the reader
    trusts them) — and some triangles are tiny, so that their cross product is_zero by the reader's 1e-5 epsilon: they stay
    in the mesh but add nothing to the vertex normals (base/STLReader.cpp:95-118)."""
    rng = np.random.default_rng(seed)
    v, f = mesh(n_tris, seed)
    corners = v[f.astype(np.int64)].astype(np.float32)
    n = np.cross(corners[:, 1] - corners[:, 0], corners[:, 2] - corners[:, 0]).astype(np.float64)
    length = np.linalg.norm(n, axis=1, keepdims=True)
    stored = np.where(length > 0, n / np.maximum(length, 1e-300), 0.0).astype(np.float32)
    kind = rng.integers(0, 4, len(f))
    stored[kind == 1] = 0.0                                                   # no normal in the file
    stored[kind == 2] = rng.normal(size=(int((kind == 2).sum()), 3)).astype(np.float32) * np.float32(3.0)
    tiny = rng.random(len(f)) < 0.05                                          # edges ~1e-3: cross ~1e-6 "is zero"
    centre = corners[tiny].mean(axis=1, keepdims=True)
    corners[tiny] = (centre + (corners[tiny] - centre) * np.float32(0.02)).astype(np.float32)
    stored[tiny & (kind != 2)] = 0.0
    return corners, stored


def write_stl(path, corners: np.ndarray, normals: np.ndarray) -> None:
    rec = np.zeros(len(corners), dtype=[("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")])
    rec["n"], rec["v"] = normals, corners
    with open(path, "wb") as fh:
        fh.write(b"binary stl written by tests/meshcases.py".ljust(80, b" "))
        fh.write(np.uint32(len(corners)).tobytes())
        fh.write(rec.tobytes())


def stl_index(corners: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """VertexIndexer (base/STLReader.cpp:18-36): a vertex gets the index of its first appearance; equality is exact
    (operator<=> on the three floats, math/Vector3.h:619-628)."""
    seen, verts = {}, []
    faces = np.zeros((len(corners), 3), dtype=np.uint32)
    for t, tri in enumerate(corners):
        for k in range(3):
            key = (float(tri[k, 0]), float(tri[k, 1]), float(tri[k, 2]))
            i = seen.get(key)
            if i is None:
                i = seen[key] = len(verts)
                verts.append(tri[k])
            faces[t, k] = i
    return np.asarray(verts, dtype=np.float32), faces
