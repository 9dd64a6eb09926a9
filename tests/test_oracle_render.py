"""The oracle's shading restatement against the REAL reference's renders (tests/golden/*.render.npz, recorded by
make_golden.py from oracle/_ref/libsp_ref.so with the reference's own mt19937_64 streams).  The oracle draws
counter-based numbers instead, so the comparison is statistical: per-pixel z-scores of the luminance means at equal
sample counts.  Stated bar: |mean z| < 0.1, < 0.5 % of pixels beyond 4 sigma, image mean within 5 standard errors."""
import numpy as np
import pytest

from conftest import GOLDEN
from simplepath_b200 import rsequence
from simplepath_b200.capi import INTEGRATORS, Partition
from simplepath_b200.flat import FlatSceneData


def lum(c):
    return 0.2126 * c[..., 0] + 0.7152 * c[..., 1] + 0.0722 * c[..., 2]


CASES = [("g_spheres", "iterative_rrnee"), ("g_spheres_ibl", "direct_lighting"), ("g_example", "iterative_rrnee"),
         ("g_example", "brute_force_iterative_rr"), ("g_bunny", "iterative_rrnee"), ("g_elf", "direct_lighting"), ("g_spheres", "whitted"),
         ("g_bunny", "whitted")]


@pytest.mark.parametrize("name,integrator", CASES)
def test_oracle_render_matches_reference_statistics(oracle_port, name, integrator):
    flat = FlatSceneData.load(GOLDEN / f"{name}.flat.npz")
    gold = np.load(GOLDEN / f"{name}.render.npz")
    spp = int(gold[f"{integrator}.spp"])
    jitter = rsequence.jitter_table(spp)
    part = Partition(0, 1, 0, spp, spp, INTEGRATORS[integrator], 424242)
    rgb, sq, st = oracle_port.render(flat.pointer(), jitter, part)
    assert st["paths"] == flat.width * flat.height * spp
    mean = lum(rgb / spp)
    var = np.maximum(sq / spp - mean ** 2, 0.0) * spp / (spp - 1)
    ref_mean, ref_var = gold[f"{integrator}.lum_mean"], gold[f"{integrator}.lum_var"]
    se = np.sqrt(var / spp + ref_var / spp) + 1e-4
    z = (mean - ref_mean) / se
    assert abs(z.mean()) < 0.1
    assert (np.abs(z) > 4).mean() < 0.005
    se_img = np.sqrt((se ** 2).sum()) / se.size
    assert abs(mean.mean() - ref_mean.mean()) < 5.0 * se_img
    # the reference's RGB means, channel by channel, on the whole image
    for ch in range(3):
        a, b = (rgb[..., ch] / spp).mean(), gold[f"{integrator}.rgb"][..., ch].mean()
        assert abs(a - b) < 0.03 * b + 1e-3


def test_rng_contract_known_answers(oracle_port):
    """The contract's counter layout is (block, stream, seed_lo, seed_hi) with key (pixel, sample): all zeros is the
    Random123 known-answer vector 6627e8d5 e169c58d bc57ac4c 9b00dbd8."""
    a = oracle_port.rng4(0, 0, 0, 0, 0)
    b = oracle_port.rng4(0, 0, 0, 0, 1)
    c = oracle_port.rng4(0, 1, 0, 0, 0)
    d = oracle_port.rng4(1, 0, 0, 0, 0)
    e = oracle_port.rng4(0, 0, 0, 1, 0)
    f = oracle_port.rng4(0, 0, 1, 0, 0)
    assert all(((x >= 0) & (x < 1)).all() for x in (a, b, c, d, e, f))
    assert len({x.tobytes() for x in (a, b, c, d, e, f)}) == 6
    # Philox4x32-10 known answer (Random123 kat_vectors): counter = key = 0
    import ctypes as C
    want = np.array([0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8], dtype=np.uint64)
    assert np.array_equal((a * 2.0 ** 24).astype(np.uint64), want >> 8)
    # uniformity over many blocks
    u = np.array([oracle_port.rng4(7, p, 3, 5, k) for p in range(64) for k in range(64)])
    assert abs(u.mean() - 0.5) < 0.01 and abs(u.var() - 1.0 / 12.0) < 0.005
