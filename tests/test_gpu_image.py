"""spcu_pack_image / spcu_render_image (include/spcu.h): the reference's output side on the device, through the C-ABI.
PFM payload: bit-exact.  PPM numbers: the sRGB curve goes through powf, whose CUDA and glibc implementations may differ in
the last bits, so a number may differ by ONE code value where 255.99 * srgb lands within a rounding error of an integer:
tolerance |difference| <= 1 on at most 0.5 % of the channels, everything else equal."""
import numpy as np
import pytest

from conftest import GOLDEN
from test_oracle_image import golden
import imagecases

pytestmark = pytest.mark.gpu


def close_numbers(got, want):
    d = np.abs(got.astype(np.int64) - want.astype(np.int64))
    assert d.max() <= 1, int(d.max())
    assert (d != 0).mean() <= 0.005, float((d != 0).mean())


def test_pack_matches_reference_files(ctx):
    from simplepath_b200 import capi
    sums, spp, pfm, ppm = golden()
    assert ctx.pack_image(sums, spp, capi.IMAGE_PFM).tobytes() == pfm.tobytes()
    close_numbers(ctx.pack_image(sums, spp, capi.IMAGE_PPM), np.minimum(ppm, 65535))


@pytest.mark.parametrize("shape", [(1, 1), (7, 3), (1920, 1080)])
def test_pack_matches_oracle(ctx, oracle_port, shape):
    from simplepath_b200 import capi
    sums, spp = imagecases.sums(seed=shape[0], w=shape[0], h=max(shape[1], 8), spp=64)
    sums = sums[:shape[1]]
    assert ctx.pack_image(sums, spp, capi.IMAGE_PFM).tobytes() == oracle_port.pack_image(sums, spp, 0).tobytes()
    close_numbers(ctx.pack_image(sums, spp, capi.IMAGE_PPM), oracle_port.pack_image(sums, spp, 1))


@pytest.mark.parametrize("name", ["g_example", "g_bunny"])
def test_render_image_is_render_frame_packed(ctx, oracle_port, name):
    """Same seed, same partition: the packed render equals the oracle's packing of the sums spcu_render_frame returns."""
    from simplepath_b200 import capi
    from simplepath_b200.flat import FlatSceneData
    flat = FlatSceneData.load(GOLDEN / f"{name}.flat.npz")
    vec = np.load(GOLDEN / f"{name}.vectors.npz")
    ctx.upload_scene(flat.pointer(), vec["jitter"], keepalive=flat)
    spp = int(vec["jitter"].shape[0])
    part = ctx.partition(spp=spp, seed=5)
    rgb, _, stats = ctx.render_frame(part, want_sumsq=False)
    pfm, st2 = ctx.render_image(part, capi.IMAGE_PFM)
    assert st2["paths"] == stats["paths"]
    assert pfm.tobytes() == oracle_port.pack_image(rgb, spp, 0).tobytes()
    ppm, _ = ctx.render_image(part, capi.IMAGE_PPM)
    close_numbers(ppm, oracle_port.pack_image(rgb, spp, 1))
    with pytest.raises(capi.SpcuError):
        ctx.pack_image(rgb, 0, capi.IMAGE_PFM)
