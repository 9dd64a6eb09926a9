"""The drop-in driver (reference main.cpp + the hunks of INTEGRATION.md) from scene file to image, with the geometry BVH
built by the reference on the host (SPCU_BUILD_ON_DEVICE=0) and left to the device (default): wall time of the whole process.

    python profiles/dropin_build_probe.py [scene ...]      # default: c3_bunny c4_elf c5_lucy
"""
import json
import os
import subprocess
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from simplepath_b200 import host, scenes  # noqa: E402

for name in (sys.argv[1:] or ["c3_bunny", "c4_elf", "c5_lucy"]):
    sp = scenes.ensure(name)
    row = {"scene": name}
    for mode in ("0", "1"):
        env = dict(os.environ, SPCU_BUILD_ON_DEVICE=mode, SPCU_SEED="3")
        t0 = time.time()
        p = subprocess.run([str(host.DRIVER), "--samples", "1", "--integrator", "cuda", sp.name], cwd=sp.parent, env=env,
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
        wall = time.time() - t0
        lines = [l for l in p.stdout.splitlines() if "CudaIntegrator" in l or "Elapsed" in l or "rror" in l]
        row["host_built" if mode == "0" else "device_built"] = {"rc": p.returncode, "wall_s": round(wall, 2), "log": lines[-3:]}
    print(json.dumps(row), flush=True)
