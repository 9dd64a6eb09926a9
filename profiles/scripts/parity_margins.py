#!/usr/bin/env python
"""How far inside their stated bounds do the per-pixel parity checks of tests/test_gpu_render.py::test_per_pixel_vs_oracle sit?
Same computation as the test, over every (scene, pipeline, integrator), printing the measured figures instead of asserting:
share of pixels outside 2e-3 * (1 + |oracle|), relative difference of the image's mean luminance, relative differences of the
ray / shade-call counters.   python profiles/scripts/parity_margins.py > profiles/rNN_parity_margins.jsonl"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from oracle import port  # noqa: E402  (test infrastructure: this script is a checker, like the tests)
from simplepath_b200 import capi  # noqa: E402
from simplepath_b200.flat import FlatSceneData  # noqa: E402

SCENES = ["g_spheres", "g_spheres_ibl", "g_example", "g_bunny", "g_elf", "g_chain", "g_lights"]
INTEGRATORS = ["iterative_rrnee", "brute_force_iterative_rr", "direct_lighting", "whitted"]
PIPELINES = {"smwave": capi.PIPELINE_SMWAVE, "paths": capi.PIPELINE_PATHS, "wavefront": capi.PIPELINE_WAVEFRONT}


def lum(c):
    return 0.2126 * c[..., 0] + 0.7152 * c[..., 1] + 0.0722 * c[..., 2]


def main():
    port.lib()
    ctx = capi.Context(0)
    golden = ROOT / "tests" / "golden"
    worst = {}
    for name in SCENES:
        flat = FlatSceneData.load(golden / f"{name}.flat.npz")
        jitter = np.load(golden / f"{name}.vectors.npz")["jitter"]
        for integrator in INTEGRATORS:
            part = None
            want = None
            for pname, pipeline in PIPELINES.items():
                ctx.set_option(capi.OPT_PIPELINE, pipeline)
                ctx.set_wavefront_size(0)
                try:
                    ctx.upload_scene(flat.pointer(), jitter, keepalive=flat)
                    part = ctx.partition(spp=jitter.shape[0], integrator=integrator, seed=20261018)
                    rgb, _, st = ctx.render(part)
                except Exception as e:  # a pipeline that does not take the scene (the test skips it too)
                    print(json.dumps({"scene": name, "integrator": integrator, "pipeline": pname, "skipped": str(e)[:80]}))
                    continue
                if want is None:
                    want, _, want_st = port.render(flat.pointer(), jitter, part)
                tol = 2e-3 * (1.0 + np.abs(want))
                bad = float((np.abs(rgb - want) > tol).any(axis=-1).mean())
                rec = {"scene": name, "integrator": integrator, "pipeline": pname, "pixels_outside_tolerance": bad,
                       "mean_luminance_rel_diff": float(abs(lum(rgb).mean() - lum(want).mean()) / max(lum(want).mean(), 1e-30))}
                for key in ("rays_closest", "rays_lights", "rays_any", "shade_calls"):
                    rec[key + "_rel_diff"] = float(abs(st[key] - want_st[key]) / max(want_st[key], 1))
                print(json.dumps(rec), flush=True)
                for k, v in rec.items():
                    if isinstance(v, float):
                        kk = k if k != "rays_any_rel_diff" or integrator not in ("direct_lighting", "whitted") else "rays_any_rel_diff_direct_whitted"
                        worst[kk] = max(worst.get(kk, 0.0), v)
    print(json.dumps({"worst_over_all": worst}))
    ctx.close()


if __name__ == "__main__":
    main()
