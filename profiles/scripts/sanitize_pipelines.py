#!/usr/bin/env python
"""Render one golden scene with one pipeline (for compute-sanitizer runs):
    [SPCU_SAN_WAVEFRONT=n] sanitize_pipelines.py SCENE PIPELINE [integrator [exact|ordered [extend|shadow]]]"""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from simplepath_b200 import capi
from simplepath_b200.flat import FlatSceneData

name, pipeline = sys.argv[1], sys.argv[2]
integrator = sys.argv[3] if len(sys.argv) > 3 else "iterative_rrnee"
flat = FlatSceneData.load(ROOT / "tests" / "golden" / f"{name}.flat.npz")
vec = np.load(ROOT / "tests" / "golden" / f"{name}.vectors.npz")
ctx = capi.Context(0)
ctx.set_option(capi.OPT_PIPELINE, {"smwave": capi.PIPELINE_SMWAVE, "paths": capi.PIPELINE_PATHS, "wavefront": capi.PIPELINE_WAVEFRONT}[pipeline])
if len(sys.argv) > 4:
    ctx.set_option(capi.OPT_TRAVERSAL, capi.TRAVERSAL_EXACT if sys.argv[4] == "exact" else capi.TRAVERSAL_ORDERED)
ctx.upload_scene(flat.pointer(), vec["jitter"], keepalive=flat)
if len(sys.argv) > 5:   # only the traversal stages, on the golden camera rays
    rays = np.ascontiguousarray(vec["camera.rays"]).view(capi.RAY_DTYPE).reshape(-1)
    if sys.argv[5] == "extend":
        hits, _ = ctx.extend_batch(rays, capi.TRAVERSAL_EXACT if sys.argv[4] == "exact" else capi.TRAVERSAL_ORDERED)
        print(name, "extend_batch ok", int((hits["id"] >= 0).sum()))
    else:
        print(name, "shadow_batch ok", int(ctx.shadow_batch(rays).sum()))
    sys.exit(0)
import os
if os.environ.get("SPCU_SAN_WAVEFRONT"):   # small batches: several of them in flight on their own streams (SPCU_OPT_BATCH_LANES)
    ctx.set_wavefront_size(int(os.environ["SPCU_SAN_WAVEFRONT"]))
rgb, _, st = ctx.render(ctx.partition(integrator=integrator, seed=77))
print(name, pipeline, integrator, "ok", float(rgb.mean()), st["paths"])
