"""Deterministic ray batches for the traversal parity tests (SURVEY.md §8c "fixtures the build must create")."""
from __future__ import annotations

import numpy as np

from simplepath_b200.capi import RAY_DTYPE

FLT_MAX = np.float32(np.finfo(np.float32).max)


def scene_bounds(flat_data) -> tuple[np.ndarray, np.ndarray]:
    """Loose bounds of the bounded geometry: union of the BVH child boxes, else a default box."""
    nodes = flat_data.arrays["geom_nodes"]
    if nodes.shape[0] == 0:
        return np.array([-5, -5, -5], np.float32), np.array([5, 5, 5], np.float32)
    box = nodes.view(np.float32).reshape(-1, 16)[:, :12].reshape(-1, 2, 6)
    lo = box[:, :, :3].reshape(-1, 3).min(0)
    hi = box[:, :, 3:].reshape(-1, 3).max(0)
    pad = 0.25 * (hi - lo)
    return (lo - pad).astype(np.float32), (hi + pad).astype(np.float32)


def _pack(o, d, t_min, t_max) -> np.ndarray:
    n = o.shape[0]
    r = np.empty(n, dtype=RAY_DTYPE)
    r["o"] = o.astype(np.float32)
    r["d"] = d.astype(np.float32)
    r["t_min"] = np.broadcast_to(np.float32(t_min), (n,))
    r["t_max"] = np.broadcast_to(np.asarray(t_max, dtype=np.float32), (n,))
    return r


def _unit(v):
    v = v.astype(np.float64)
    return (v / np.linalg.norm(v, axis=1, keepdims=True)).astype(np.float32)


def random_rays(flat_data, n: int, seed: int = 12345) -> np.ndarray:
    rng = np.random.default_rng(seed)
    lo, hi = scene_bounds(flat_data)
    o = rng.uniform(lo, hi, size=(n, 3))
    d = _unit(rng.normal(size=(n, 3)))
    return _pack(o, d, 1e-3, FLT_MAX)


def segment_rays(flat_data, n: int, seed: int = 777) -> np.ndarray:
    """Shadow-ray style: finite t_max, grazing-dependent t_min."""
    rng = np.random.default_rng(seed)
    lo, hi = scene_bounds(flat_data)
    a = rng.uniform(lo, hi, size=(n, 3))
    b = rng.uniform(lo, hi, size=(n, 3))
    dist = np.linalg.norm(b - a, axis=1)
    d = _unit(b - a)
    t_max = (dist * rng.uniform(0.2, 1.2, size=n)).astype(np.float32)
    t_min = (1e-3 / np.maximum(rng.uniform(0.0, 1.0, size=n), 1e-3)).astype(np.float32)
    r = _pack(a, d, 0.0, t_max)
    r["t_min"] = t_min
    return r


def axis_rays(flat_data, n: int, seed: int = 4242) -> np.ndarray:
    """Directions with one or two exactly-zero components: 1/0 = inf and 0*inf = NaN in the slab test
    (math/BBox.h:128-142)."""
    rng = np.random.default_rng(seed)
    lo, hi = scene_bounds(flat_data)
    o = rng.uniform(lo, hi, size=(n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3))
    kind = rng.integers(0, 6, size=n)
    for k in range(3):
        d[kind == k, k] = 0.0                       # one zero component
        two = kind == 3 + k
        d[two] = 0.0
        d[two, k] = rng.choice([-1.0, 1.0], size=int(two.sum()))
    # half of the origins sit exactly on a node-box plane so that (lo - o) == 0 meets inv == inf
    nodes = flat_data.arrays["geom_nodes"]
    if nodes.shape[0]:
        box = nodes.view(np.float32).reshape(-1, 16)[:, :12]
        pick = rng.integers(0, box.shape[0], size=n)
        comp = rng.integers(0, 12, size=n)
        axis = np.array([0, 1, 2, 0, 1, 2, 0, 1, 2, 0, 1, 2])[comp]
        snap = rng.random(n) < 0.5
        o[snap, axis[snap]] = box[pick[snap], comp[snap]]
    return _pack(o, _unit(d), 1e-3, FLT_MAX)


def grazing_rays(flat_data, n: int, seed: int = 999) -> np.ndarray:
    """Rays aimed exactly at triangle vertices and edge midpoints: beta/gamma on their limits, shared edges hit at
    equal t by two triangles (the `t > t_max` rule makes the later primitive win, shapes/Triangle.h:143)."""
    rng = np.random.default_rng(seed)
    prims = flat_data.arrays["geom_prims"].view(np.float32).reshape(-1, 12)
    meta = flat_data.arrays["geom_meta"].view(np.uint32).reshape(-1)
    tris = prims[(meta & 3) == 0]
    if tris.shape[0] == 0:
        return random_rays(flat_data, n, seed)
    lo, hi = scene_bounds(flat_data)
    t = tris[rng.integers(0, tris.shape[0], size=n)]
    v = np.stack([t[:, 0:3], t[:, 4:7], t[:, 8:11]], axis=1)             # [n, 3, 3]
    k = rng.integers(0, 3, size=n)
    a = v[np.arange(n), k]
    b = v[np.arange(n), (k + 1) % 3]
    mid = rng.random(n) < 0.5
    target = np.where(mid[:, None], (0.5 * (a + b)).astype(np.float32), a)
    o = rng.uniform(lo, hi, size=(n, 3)).astype(np.float32)
    d = _unit(target - o)
    return _pack(o, d, 1e-3, FLT_MAX)


def cone_rays(n: int, seed: int = 31337) -> np.ndarray:
    """For the chain mesh of simplepath_b200.scenes.chain_mesh (a self-similar cone of triangles around the +x axis with
    its apex at the origin): rays that cross MANY of its nested node boxes — from outside towards the apex (the
    reference-order walk then leaves a pending right child at every level: the deepest traversal stack) and from next to
    the apex outwards, starting between two triangles at a random level (the ordered walk's deepest stack)."""
    rng = np.random.default_rng(seed)
    half = n // 2
    m = rng.uniform(-0.12, 0.12, size=(n, 2))                   # slopes inside the cone (its half-width is 0.2 x)
    o = np.empty((n, 3))
    d = np.empty((n, 3))
    x_out = rng.uniform(1.2, 3.0, size=half)
    o[:half] = np.stack([x_out, m[:half, 0] * x_out, m[:half, 1] * x_out], axis=1)
    aim = rng.normal(scale=1e-3, size=(half, 3)) * rng.choice([0.0, 1.0], size=(half, 1))   # half of them exactly at the apex
    d[:half] = aim - o[:half]
    k = rng.integers(0, 57, size=n - half)
    x_in = 0.45 ** k * rng.uniform(0.5, 0.95, size=n - half)    # between triangle k and k + 1
    o[half:] = np.stack([x_in, m[half:, 0] * x_in, m[half:, 1] * x_in], axis=1)
    m2 = rng.uniform(-0.12, 0.12, size=(n - half, 2))
    d[half:] = np.stack([np.ones(n - half), m2[:, 0], m2[:, 1]], axis=1)
    return _pack(o.astype(np.float32), _unit(d), 0.0, FLT_MAX)


def all_batches(flat_data, n_each: int) -> dict[str, np.ndarray]:
    return {
        "random": random_rays(flat_data, n_each),
        "segments": segment_rays(flat_data, n_each),
        "axis": axis_rays(flat_data, n_each),
        "grazing": grazing_rays(flat_data, n_each),
    }
