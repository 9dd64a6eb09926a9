"""The drop-in itself: the reference's own main.cpp with the three hunks of INTEGRATION.md (SimplePathCuda), selecting
the CUDA backend through `--integrator cuda`, writes the same image as the Python mirror of the host side."""
import os
import subprocess

import numpy as np
import pytest

from simplepath_b200 import capi, host, scenes

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(not host.DRIVER.exists() or not host.available(), reason="host plugin not built (needs the reference sources)")
@pytest.mark.parametrize("name,flag,integrator", [("g_bunny", "cuda", "iterative_rrnee"),
                                                  ("g_spheres_ibl", "cuda_direct_lighting", "direct_lighting")])
def test_driver_writes_the_same_image(ctx, tmp_path, name, flag, integrator):
    sp = scenes.ensure(name, tmp_path)
    spp = 4
    env = dict(os.environ, SPCU_SEED="11")
    proc = subprocess.run([str(host.DRIVER), "--samples", str(spp), "--integrator", flag, sp.name], cwd=tmp_path, env=env,
                          stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
    assert proc.returncode == 0, proc.stdout[-2000:]
    assert "Elapsed time" in proc.stdout
    img = scenes.read_pfm(tmp_path / "image.pfm")

    flat = host.parse(sp)
    ctx.set_wavefront_size(0)
    ctx.upload_scene(flat.pointer(), host.jitter(spp), keepalive=flat)
    rgb, _, _ = ctx.render(ctx.partition(spp=spp, integrator=integrator, seed=11), want_sumsq=False)
    want = rgb / np.float32(spp)
    assert img.shape == want.shape
    assert img.tobytes() == want.astype(np.float32).tobytes()


@pytest.mark.skipif(not host.DRIVER.exists(), reason="host plugin not built")
def test_driver_reports_errors_like_the_reference(tmp_path):
    proc = subprocess.run([str(host.DRIVER), "--integrator", "cuda_mandelbrot", "nothing.sp"], cwd=tmp_path,
                          stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=60)
    assert proc.returncode != 0
    assert "no device path" in proc.stdout


@pytest.mark.skipif(not host.DRIVER.exists() or not host.available(), reason="host plugin not built (needs the reference sources)")
def test_driver_over_two_devices_is_bit_identical(tmp_path):
    """SPCU_DEVICES=2: tiles interleaved over two GPUs touch disjoint pixels, so the image equals the one-GPU image."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    sp = scenes.ensure("g_bunny", tmp_path)
    images = []
    for n in (1, 2):
        env = dict(os.environ, SPCU_SEED="5", SPCU_DEVICES=str(n))
        proc = subprocess.run([str(host.DRIVER), "--samples", "4", "--integrator", "cuda", sp.name], cwd=tmp_path, env=env,
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=180)
        assert proc.returncode == 0, proc.stdout[-2000:]
        images.append(scenes.read_pfm(tmp_path / "image.pfm"))
    assert images[0].tobytes() == images[1].tobytes()


@pytest.mark.skipif(not host.DRIVER.exists() or not host.available(), reason="host plugin not built (needs the reference sources)")
def test_driver_with_the_tree_built_on_the_device(tmp_path):
    """SPCU_BUILD_ON_DEVICE=1: the plugin hands the geometry over unbuilt (spcu_upload_scene_build); the device must arrive at
    the reference's own tree (checked inside the plugin) and the image is the same, bit for bit."""
    sp = scenes.ensure("g_bunny", tmp_path)
    images = []
    for build in ("0", "1"):
        env = dict(os.environ, SPCU_SEED="5", SPCU_BUILD_ON_DEVICE=build)
        proc = subprocess.run([str(host.DRIVER), "--samples", "4", "--integrator", "cuda", sp.name], cwd=tmp_path, env=env,
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=180)
        assert proc.returncode == 0, proc.stdout[-2000:]
        assert ("acceleration structure built on device" in proc.stdout) == (build == "1")
        images.append(scenes.read_pfm(tmp_path / "image.pfm"))
    assert images[0].tobytes() == images[1].tobytes()
