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


def _clone(flat):
    from simplepath_b200.flat import FlatSceneData
    import copy
    return FlatSceneData(copy.deepcopy(flat.head), {k: v.copy() for k, v in flat.arrays.items()})


@pytest.mark.parametrize("name", ["g_bunny", "g_elf"])
def test_upload_scene_build(ctx, oracle_port, name):
    """spcu_upload_scene_build: primitives in a shuffled pre-construction order; bounds, tree and leaf-order gather on the
    device.  Checked against the all-CPU pipeline (oracle bounds -> oracle construction -> oracle traversal), and, from
    the reference's own leaf order, against the golden answers and a bitwise-equal render."""
    from simplepath_b200 import capi
    from simplepath_b200.flat import FlatSceneData
    flat = FlatSceneData.load(GOLDEN / f"{name}.flat.npz")
    vec = np.load(GOLDEN / f"{name}.vectors.npz")
    nu, n = flat.head["geom"]["n_unbounded"], flat.head["geom"]["n_prims"] - flat.head["geom"]["n_unbounded"]
    rays = np.concatenate([np.ascontiguousarray(vec[f"{b}.rays"]).view(capi.RAY_DTYPE).reshape(-1)
                           for b in ("camera", "random", "segments", "axis", "grazing")])

    perm = np.random.default_rng(9).permutation(n)
    shuffled = _clone(flat)
    for k in ("geom_prims", "geom_shade", "geom_meta"):
        shuffled.arrays[k][nu:] = flat.arrays[k][nu + perm]
    tris = shuffled.arrays["geom_prims"].view(np.float32).reshape(-1, 12)[nu:]
    built = oracle_port.build_bvh(oracle_port.triangle_bounds(tris), None, nu)
    cpu = _clone(shuffled)
    for k in ("geom_prims", "geom_shade", "geom_meta"):
        cpu.arrays[k][nu:] = shuffled.arrays[k][nu + built["order"]]
    cpu.arrays["geom_nodes"] = built["nodes"].view(np.uint8).reshape(-1, 64).copy()
    cpu.head["geom"] = built["head"]
    want = oracle_port.trace_closest(cpu.pointer(), rays)
    want_any = oracle_port.trace_any(cpu.pointer(), rays)

    order, head = ctx.upload_scene_build(shuffled.pointer(), vec["jitter"], keepalive=shuffled)
    assert head == built["head"]
    assert np.array_equal(order, built["order"])
    got = ctx.trace_closest(rays)
    assert np.array_equal(got["id"], want["id"])
    assert got["t"].tobytes() == want["t"].tobytes()
    assert np.array_equal(ctx.trace_any(rays), want_any)
    # same triangles as the reference's tree found (IDs differ: another order, another tree)
    cam = np.ascontiguousarray(vec["camera.rays"]).view(capi.RAY_DTYPE).reshape(-1)
    hit = ctx.trace_closest(cam)
    ref_id = vec["camera.closest_id"]
    assert np.array_equal(hit["id"] >= 0, ref_id >= 0)
    bounded = hit["id"] >= nu
    original = nu + perm[order[hit["id"][bounded] - nu]]          # device ID -> shuffled index -> golden leaf position
    same = original == ref_id[bounded]
    assert same.mean() > 0.999, same.mean()                        # equal-distance ties on shared edges may pick the neighbour
    assert (hit["t"][bounded][same].view(np.uint32) == vec["camera.closest_t"][bounded][same].view(np.uint32)).all()

    # from the reference's leaf order the construction is a fixed point: the golden answers, and the same image bit for bit
    ctx.upload_scene(flat.pointer(), vec["jitter"], keepalive=flat)
    part = ctx.partition(spp=int(vec["jitter"].shape[0]), seed=3)
    rgb_flat, _, _ = ctx.render(part)
    order, head = ctx.upload_scene_build(flat.pointer(), vec["jitter"], keepalive=flat)
    assert head == flat.head["geom"] and np.array_equal(order, np.arange(n))
    hit = ctx.trace_closest(cam)
    assert np.array_equal(hit["id"], ref_id) and hit["t"].tobytes() == vec["camera.closest_t"].tobytes()
    rgb_built, _, _ = ctx.render(part)
    assert rgb_built.tobytes() == rgb_flat.tobytes()


def test_upload_scene_build_analytic_and_errors(ctx):
    """Scenes without a mesh: spheres need caller-supplied bounds; four spheres make one root leaf (material_spheres)."""
    from simplepath_b200 import capi
    from simplepath_b200.flat import FlatSceneData
    flat = FlatSceneData.load(GOLDEN / "g_spheres.flat.npz")
    vec = np.load(GOLDEN / "g_spheres.vectors.npz")
    with pytest.raises(capi.SpcuError):
        ctx.upload_scene_build(flat.pointer(), vec["jitter"], keepalive=flat)     # spheres, no bounds
    g = flat.head["geom"]
    n = g["n_prims"] - g["n_unbounded"]
    bounds = bvhcases.boxes(np.random.default_rng(2), n)                          # <= 4 primitives: any bounds give a leaf
    order, head = ctx.upload_scene_build(flat.pointer(), vec["jitter"], bounds=bounds, keepalive=flat)
    assert head == g and np.array_equal(order, np.arange(n))
    cam = np.ascontiguousarray(vec["camera.rays"]).view(capi.RAY_DTYPE).reshape(-1)
    hit = ctx.trace_closest(cam)
    assert np.array_equal(hit["id"], vec["camera.closest_id"]) and hit["t"].tobytes() == vec["camera.closest_t"].tobytes()


def test_upload_scene_build_with_spheres_inside_the_tree(ctx, oracle_port):
    """Forty spheres (copies of material_spheres' four, moved apart) and their caller-supplied bounds: a real tree whose leaves
    carry the mixed flag.  Same tree from the oracle, same hits from the oracle's traversal of it."""
    from simplepath_b200 import capi
    from simplepath_b200.flat import FlatSceneData
    flat = FlatSceneData.load(GOLDEN / "g_spheres.flat.npz")
    vec = np.load(GOLDEN / "g_spheres.vectors.npz")
    g = flat.head["geom"]
    nu, n0 = g["n_unbounded"], g["n_prims"] - g["n_unbounded"]
    rng = np.random.default_rng(6)
    prims0 = flat.arrays["geom_prims"].view(np.float32).reshape(-1, 12)
    recs, bounds = {k: [flat.arrays[k][:nu]] for k in ("geom_prims", "geom_shade", "geom_meta")}, []
    for copy in range(10):
        shift = (rng.uniform(-6, 6, 3) * np.array([1.0, 0.2, 1.0])).astype(np.float64) if copy else np.zeros(3)
        for i in range(nu, nu + n0):
            w2o = prims0[i].astype(np.float64)                      # c0.xyz c1.xyz c2.xyz affine.xyz
            lin = w2o[:9].reshape(3, 3).T
            moved = prims0[i].copy()
            moved[9:12] = (w2o[9:12] - lin @ shift).astype(np.float32)
            o2w = np.linalg.inv(lin)
            centre = -o2w @ moved[9:12].astype(np.float64)
            reach = np.abs(o2w).sum(axis=1) * 1.001                 # the unit sphere's box through object_to_world, padded
            bounds.append(np.concatenate([centre - reach, centre + reach]).astype(np.float32))
            recs["geom_prims"].append(moved.view(np.uint8).reshape(1, 48))
            recs["geom_shade"].append(flat.arrays["geom_shade"][i:i + 1])
            recs["geom_meta"].append(flat.arrays["geom_meta"][i:i + 1])
    s = _clone(flat)
    for k in recs:
        s.arrays[k] = np.concatenate(recs[k])
    n = 10 * n0
    s.head["geom"] = dict(g, n_prims=nu + n, n_nodes=0, root=~nu, root_count=0, max_depth=0)
    bounds = np.stack(bounds)
    kinds = s.arrays["geom_meta"].view(np.uint32).reshape(-1)[nu:] & 3
    built = oracle_port.build_bvh(bounds, (kinds != 0).astype(np.uint8), nu)
    assert built["head"]["n_nodes"] > 3
    leaf = built["nodes"]["child"] < 0
    assert ((built["nodes"]["count"][leaf] & 0x80000000) != 0).all()          # every leaf holds spheres
    order, head = ctx.upload_scene_build(s.pointer(), vec["jitter"], bounds=bounds, keepalive=s)
    assert head == built["head"] and np.array_equal(order, built["order"])
    cpu = _clone(s)
    for k in ("geom_prims", "geom_shade", "geom_meta"):
        cpu.arrays[k][nu:] = s.arrays[k][nu + built["order"]]
    cpu.arrays["geom_nodes"] = built["nodes"].view(np.uint8).reshape(-1, 64).copy()
    cpu.head["geom"] = built["head"]
    rays = np.concatenate([np.ascontiguousarray(vec[f"{b}.rays"]).view(capi.RAY_DTYPE).reshape(-1) for b in ("camera", "random", "segments")])
    want, got = oracle_port.trace_closest(cpu.pointer(), rays), ctx.trace_closest(rays)
    assert (want["id"] >= nu).sum() > 1000
    assert np.array_equal(got["id"], want["id"]) and got["t"].tobytes() == want["t"].tobytes()
    assert np.array_equal(ctx.trace_any(rays), oracle_port.trace_any(cpu.pointer(), rays))
