"""spcu_build_bvh / spcu_triangle_bounds (include/spcu.h) — BVHAccelerator::construct on the device — through the C-ABI:
bit-identical nodes, primitive order and accel header against trees the reference built (tests/golden/bvh_build.npz), the
oracle on fresh inputs, the golden scenes' own BVHs, and at full size through properties that do not need a second build."""
import numpy as np
import pytest

import bvhcases
from test_oracle_build import CASES, golden_build, scene_bvh
from conftest import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(CASES))
def test_build_matches_reference_trees(ctx, name):
    bounds, non_tri, first_id = CASES[name]
    got = ctx.build_bvh(bounds, non_tri, first_id)
    want = golden_build(name)
    assert not bvhcases.same(got, want), bvhcases.same(got, want)


@pytest.mark.parametrize("n", [31, 32, 33, 2047, 2048, 2049, 100_000, 1_000_000])
def test_build_matches_oracle_on_fresh_boxes(ctx, oracle_port, n):
    rng = np.random.default_rng(n)
    b = bvhcases.boxes(rng, n, spread=float(rng.uniform(0.1, 100)), size=float(rng.uniform(1e-4, 0.2)))
    nt = (rng.random(n) < 0.02).astype(np.uint8)
    got, want = ctx.build_bvh(b, nt, 11), oracle_port.build_bvh(b, nt, 11)
    assert not bvhcases.same(got, want), bvhcases.same(got, want)


@pytest.mark.parametrize("name", ["g_bunny", "g_elf"])
def test_triangle_bounds_and_scene_fixed_point(ctx, name):
    tris, nodes, head = scene_bvh(name)
    bounds = ctx.triangle_bounds(tris)
    assert bounds.tobytes() == np.load(GOLDEN / f"{name}.bounds.npz")["bounds"][head["n_unbounded"]:].tobytes()
    got = ctx.build_bvh(bounds, None, head["n_unbounded"])
    assert got["head"] == head
    assert got["nodes"].tobytes() == nodes.tobytes()
    assert np.array_equal(got["order"], np.arange(len(tris)))


def test_rebuilt_tree_traces_like_the_reference_tree(ctx):
    """Upload g_bunny with the device-built BVH in place of the flattener's: identical camera-ray answers."""
    from simplepath_b200.flat import FlatSceneData
    flat = FlatSceneData.load(GOLDEN / "g_bunny.flat.npz")
    vec = np.load(GOLDEN / "g_bunny.vectors.npz")
    tris, _, head = scene_bvh("g_bunny")
    got = ctx.build_bvh(ctx.triangle_bounds(tris), None, head["n_unbounded"])
    flat.arrays["geom_nodes"] = got["nodes"].view(np.uint8).reshape(-1, 64).copy()
    flat.head["geom"] = got["head"]
    flat._struct = None
    ctx.upload_scene(flat.pointer(), vec["jitter"], keepalive=flat)
    hits = ctx.trace_closest(np.ascontiguousarray(vec["camera.rays"]).view(ctx_ray_dtype()).reshape(-1))
    assert np.array_equal(hits["id"], vec["camera.closest_id"])
    assert hits["t"].tobytes() == vec["camera.closest_t"].tobytes()


def ctx_ray_dtype():
    from simplepath_b200 import capi
    return capi.RAY_DTYPE


def test_build_error_paths(ctx):
    from simplepath_b200 import capi
    b = bvhcases.boxes(np.random.default_rng(3), 100)
    with pytest.raises(capi.SpcuError):
        ctx.build_bvh(b, None, 0, capacity=3)  # 100 boxes need more than 3 internal nodes


def test_build_full_size_properties(ctx):
    """28 M boxes (the lucy scene's primitive count is 28,055,742): every primitive appears once; every leaf holds at most
    four primitives; pre-order links are consistent; each child box is the union of its primitives' boxes."""
    n = 28_055_742
    rng = np.random.default_rng(5)
    c = rng.random((n, 3), dtype=np.float32) * np.float32(1000.0)
    h = rng.random((n, 3), dtype=np.float32) * np.float32(0.05)
    b = np.concatenate([c - h, c + h], axis=1)
    got = ctx.build_bvh(b, None, 1)
    order, nodes, head = got["order"], got["nodes"], got["head"]
    print(f"\nspcu_build_bvh: {n} boxes -> {head['n_nodes']} internal nodes, depth {head['max_depth']}, "
          f"{got['device_ms']:.1f} ms on the device")
    seen = np.zeros(n, dtype=np.uint8)
    seen[order] = 1
    assert seen.all()
    child, count = nodes["child"], nodes["count"]
    leaf = child < 0
    assert (count[leaf] & 0x7FFFFFFF).max() <= 4 and (count[leaf] & 0x7FFFFFFF).min() >= 1
    assert int((count[leaf] & 0x7FFFFFFF).sum()) == n
    # pre-order: the left child of node i, when internal, is i + 1; every internal link points forward
    left_internal = ~leaf[:, 0]
    idx = np.arange(len(nodes))
    assert np.array_equal(child[left_internal, 0], idx[left_internal] + 1)
    assert (child[~leaf[:, 1], 1] > idx[~leaf[:, 1]]).all()
    # leaf ranges tile [first_id, first_id + n) in link order
    firsts = np.sort((~child[leaf]).astype(np.int64))
    assert firsts[0] == 1 and np.all(np.diff(firsts) >= 1)
    # boxes of a sample of leaf children = union of their primitives' boxes
    lk = np.argwhere(leaf)
    for i, k in lk[rng.integers(0, len(lk), 2000)]:
        first, cnt = int(~child[i, k]) - 1, int(count[i, k] & 0x7FFFFFFF)
        pb = b[order[first:first + cnt]]
        want = np.concatenate([pb[:, :3].min(0), pb[:, 3:].max(0)])
        assert np.array_equal(nodes["box"][i, 6 * k:6 * k + 6], want)
