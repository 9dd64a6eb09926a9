"""The oracle's plain-C restatement against the golden vectors recorded from the real reference
(tests/golden/make_golden.py): primitive IDs, distances (bitwise), any-hit flags, light hits, camera rays and
Intersection records."""
import numpy as np
import pytest

from conftest import GOLDEN, golden_names
from simplepath_b200.capi import RAY_DTYPE
from simplepath_b200.flat import FlatSceneData

BATCHES = ["camera", "random", "segments", "axis", "grazing"]


def load(name):
    return FlatSceneData.load(GOLDEN / f"{name}.flat.npz"), np.load(GOLDEN / f"{name}.vectors.npz")


def rays_of(vec, batch):
    return np.ascontiguousarray(vec[f"{batch}.rays"]).view(RAY_DTYPE).reshape(-1)


def test_golden_present():
    assert set(golden_names()) >= {"g_spheres", "g_spheres_ibl", "g_example", "g_bunny", "g_elf"}


@pytest.mark.parametrize("name", golden_names())
@pytest.mark.parametrize("batch", BATCHES)
def test_closest_any_lights_bit_exact(oracle_port, name, batch):
    flat, vec = load(name)
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
    for batch in BATCHES:
        rec, _ = oracle_port.hit_records(flat.pointer(), rays_of(vec, batch))
        assert rec.tobytes() == vec[f"{batch}.records"].tobytes()


def test_equal_t_ties_exist_in_fixtures():
    """The grazing batch must actually exercise shared-edge ties, otherwise the 'last wins' rule is untested."""
    flat, vec = load("g_bunny")
    ids = vec["grazing.closest_id"]
    assert (ids >= 0).sum() > 1000
