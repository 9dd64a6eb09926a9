"""The oracle's plain-C restatement against the golden vectors recorded from the real reference
(tests/golden/make_golden.py): primitive IDs, distances (bitwise), any-hit flags, light hits, camera rays and
Intersection records."""
import numpy as np
import pytest

from conftest import GOLDEN, golden_names
from simplepath_b200.capi import RAY_DTYPE
from simplepath_b200.flat import FlatSceneData

BATCHES = ["camera", "random", "segments", "axis", "grazing", "cone"]  # "cone": g_chain only (deep traversal stacks)


def load(name):
    return FlatSceneData.load(GOLDEN / f"{name}.flat.npz"), np.load(GOLDEN / f"{name}.vectors.npz")


def rays_of(vec, batch):
    return np.ascontiguousarray(vec[f"{batch}.rays"]).view(RAY_DTYPE).reshape(-1)


def test_golden_present():
    assert set(golden_names()) >= {"g_spheres", "g_spheres_ibl", "g_example", "g_bunny", "g_elf", "g_chain", "g_lights"}


def _stack_depths(flat, rays, ordered):
    """Deepest traversal stack (pending children) each ray's walk reaches, by a plain float64 re-enactment of the two walks
    (reference order: left first, the right child waits; ordered: nearer child first).  Only used to prove that the
    fixtures reach the depths they are meant to reach — not a parity check."""
    nodes = flat.arrays["geom_nodes"].view(np.float32).reshape(-1, 16).astype(np.float64)
    child = flat.arrays["geom_nodes"].view(np.int32).reshape(-1, 16)[:, 12:14]
    out = []
    for r in rays:
        o, d = r["o"].astype(np.float64), r["d"].astype(np.float64)
        with np.errstate(divide="ignore", invalid="ignore"):
            inv = 1.0 / d

        def entry(box):
            with np.errstate(invalid="ignore"):
                t0, t1 = (box[:3] - o) * inv, (box[3:] - o) * inv
            lo, hi = np.fmin(t0, t1).max(), np.fmax(t0, t1).min()
            return lo if max(lo, float(r["t_min"])) <= hi else None
        stack, deepest = [0], 0
        while stack:
            deepest = max(deepest, len(stack))
            n = stack.pop()
            if n < 0:
                continue
            e = [entry(nodes[n, 0:6]), entry(nodes[n, 6:12])]
            kids = [k for k in (0, 1) if e[k] is not None]
            if ordered:
                kids.sort(key=lambda k: e[k])
            for k in reversed(kids):   # the first to visit goes on top
                stack.append(int(child[n, k]))
        out.append(deepest)
    return np.array(out)


def test_chain_scene_reaches_the_overflow_stack():
    """g_chain is a BVH of depth 54; its cone batch must drive BOTH walks past the 24 (exact) / 12 (ordered) stack levels
    the kernels keep in shared memory, into the thread-local overflow (csrc/trace.cuh Stack::loc, OrderedStack::loc)."""
    flat, vec = load("g_chain")
    assert flat.head["geom"]["max_depth"] >= 40
    rays = rays_of(vec, "cone")
    # the half of the batch that starts next to the apex and looks outwards: every enclosing level's far child waits
    exact = _stack_depths(flat, rays[-256:], ordered=False)
    ordered = _stack_depths(flat, rays[-256:], ordered=True)
    assert exact.max() > 40 and (exact > 24).sum() >= 32
    assert ordered.max() > 24 and (ordered > 12).sum() >= 32


def test_lights_scene_has_a_lights_bvh():
    """g_lights: Scene's lights accelerator is [environment, BVH(7 sphere lights)] with internal nodes
    (shapes/BVHAccelerator.h:45-60 NodeInternal::intersect_lights), and the batches do reach lights through it."""
    flat, vec = load("g_lights")
    la = flat.head["lights_accel"]
    assert la["n_nodes"] >= 1 and la["n_unbounded"] == 1 and la["n_prims"] == 8
    ids = np.concatenate([vec[f"{b}.lights_id"] for b in BATCHES if f"{b}.rays" in vec])
    assert len(set(ids[ids >= 1].tolist())) == 7          # every sphere light is hit by some ray


@pytest.mark.parametrize("name", golden_names())
@pytest.mark.parametrize("batch", BATCHES)
def test_closest_any_lights_bit_exact(oracle_port, name, batch):
    flat, vec = load(name)
    if f"{batch}.rays" not in vec:
        pytest.skip(f"{name} has no {batch} batch")
    rays = rays_of(vec, batch)
    hits, cnt = oracle_port.trace_closest(flat.pointer(), rays, counters=True)
    assert np.array_equal(hits["id"], vec[f"{batch}.closest_id"])
    assert hits["t"].tobytes() == vec[f"{batch}.closest_t"].tobytes()
    assert np.array_equal(cnt, vec[f"{batch}.counters"])
    assert np.array_equal(oracle_port.trace_any(flat.pointer(), rays), vec[f"{batch}.any"])
    lh = oracle_port.trace_lights(flat.pointer(), rays)
    assert np.array_equal(lh["id"], vec[f"{batch}.lights_id"])
    assert lh["t"].tobytes() == vec[f"{batch}.lights_t"].tobytes()


@pytest.mark.parametrize("name", golden_names())
def test_camera_rays_and_records(oracle_port, name):
    flat, vec = load(name)
    rays = oracle_port.generate_rays(flat.pointer(), vec["jitter"], vec["cam_pix"], vec["cam_smp"])
    assert rays.tobytes() == rays_of(vec, "camera").tobytes()
    for batch in (b for b in BATCHES if f"{b}.rays" in vec):
        rec, _ = oracle_port.hit_records(flat.pointer(), rays_of(vec, batch))
        assert rec.tobytes() == vec[f"{batch}.records"].tobytes()


def test_equal_t_ties_exist_in_fixtures():
    """The grazing batch must actually exercise shared-edge ties, otherwise the 'last wins' rule is untested."""
    flat, vec = load("g_bunny")
    ids = vec["grazing.closest_id"]
    assert (ids >= 0).sum() > 1000
