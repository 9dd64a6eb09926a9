"""Box sets for the BVH-construction parity tests (BVHAccelerator::construct, reference shapes/BVHAccelerator.h:175-209):
seeded, so the golden file (tests/golden/make_golden_bvh.py, answers from the reference's own BVHAccelerator) and the tests
see the same inputs.  Each case = (bounds [n, 6] float32 in INITIAL order, non_triangle [n] uint8 | None, first_id)."""
from __future__ import annotations

import numpy as np


def boxes(rng, n: int, spread: float = 1.0, size: float = 0.05) -> np.ndarray:
    c = rng.uniform(-spread, spread, (n, 3)).astype(np.float32)
    h = rng.uniform(0, size, (n, 3)).astype(np.float32)
    return np.concatenate([c - h, c + h], axis=1).astype(np.float32)


def signed_zeros(rng, shape) -> np.ndarray:
    return np.where(rng.random(shape) < 0.5, np.float32(-0.0), np.float32(0.0)).astype(np.float32)


def sorted_corners(b: np.ndarray) -> np.ndarray:
    return np.concatenate([np.minimum(b[:, :3], b[:, 3:]), np.maximum(b[:, :3], b[:, 3:])], axis=1).astype(np.float32)


def cases(seed: int = 20261018, big: int = 20000) -> dict:
    rng = np.random.default_rng(seed)
    out = {}
    for n in (0, 1, 2, 4, 5, 6, 9, 33, 1000):
        out[f"random_{n}"] = (boxes(rng, n), None, 0)
    out[f"random_{big}"] = (boxes(rng, big), None, 3)
    # leaf of more than four primitives: a partition with an empty side (BVHAccelerator.h:200-203)
    b = boxes(rng, 50)
    b[10:40] = b[10]
    out["duplicates"] = (b, None, 0)
    out["all_equal_17"] = (np.tile(boxes(rng, 1), (17, 1)), None, 1)
    # clustered: deep, unbalanced tree
    c = np.concatenate([boxes(rng, 400, spread=1e-3, size=1e-5), boxes(rng, 40, spread=50.0)])
    out["clustered"] = (c[rng.permutation(len(c))], None, 0)
    # sign of zero: ties in _mm_min_ps / _mm_max_ps and BBox's corner re-sort (math/BBox.h:26-30,60-64)
    b = boxes(rng, 200)
    b[:, 0] = signed_zeros(rng, 200)
    b[:, 3] = signed_zeros(rng, 200)
    out["zero_axis"] = (b, None, 0)
    out["all_zero_40"] = (signed_zeros(rng, (40, 6)), None, 0)
    for t in range(3):
        b = boxes(rng, 300)
        z = rng.random((300, 6)) < 0.3
        b[z] = signed_zeros(rng, int(z.sum()))
        out[f"mixed_zero_{t}"] = (sorted_corners(b), None, 0)
    for t in range(8):  # a prefix of [+-0, +-0] intervals followed by intervals that open the axis
        n = int(rng.integers(5, 12))
        b = signed_zeros(rng, (n, 6))
        k = int(rng.integers(0, n))
        b[k:, 3 + int(rng.integers(0, 3))] = rng.random(n - k).astype(np.float32)
        out[f"zero_prefix_{t}"] = (b, None, 0)
    # flat geometry (axis-aligned quads in the plane y = 0, like a tessellated floor) and a regular grid with many ties
    g = np.stack(np.meshgrid(np.arange(32, dtype=np.float32), np.arange(32, dtype=np.float32)), -1).reshape(-1, 2)
    flat = np.zeros((len(g), 6), dtype=np.float32)
    flat[:, 0], flat[:, 2], flat[:, 3], flat[:, 5] = g[:, 0] - 16, g[:, 1] - 16, g[:, 0] - 15, g[:, 1] - 15
    out["floor_grid"] = (flat, None, 2)
    out["floor_grid_shuffled"] = (flat[rng.permutation(len(flat))], None, 2)
    # spheres among triangles: the leaf's mixed flag
    out["mixed_kinds"] = (boxes(rng, 1000), (rng.random(1000) < 0.05).astype(np.uint8), 7)
    out["mixed_kinds_root_leaf"] = (boxes(rng, 3), np.array([0, 1, 0], dtype=np.uint8), 1)
    return out


def same(a: dict, b: dict) -> list[str]:
    """Names of the fields in which two build results differ (bitwise for floats)."""
    bad = []
    if a["head"] != b["head"]:
        bad.append(f"head {a['head']} != {b['head']}")
    if a["nodes"].tobytes() != b["nodes"].tobytes():
        bad.append("nodes")
    if not np.array_equal(a["order"], b["order"]):
        bad.append("order")
    return bad
