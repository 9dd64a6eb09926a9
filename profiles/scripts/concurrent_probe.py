#!/usr/bin/env python
"""Does the GPU do more work per second when the kernels of several batches share it?  N contexts of one process render the N
parts of a frame's samples at the same time on their own streams, against one context rendering all of them alone: the
measurement behind SPCU_OPT_BATCH_LANES (records profiles/r02x_*, r02y_*).  (Those records also hold runs with
every grid sized for 1/2 or 1/3 of the machine — "grid_divisor" — through an environment switch the library had for the
experiment: half-size grids lost, the switch is gone.)

    python profiles/scripts/concurrent_probe.py WORKLOAD SPP N_CONTEXTS [frames]
"""
import ctypes as C
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from simplepath_b200 import host  # noqa: E402
from simplepath_b200.capi import FlatScene, Partition, Stats, INTEGRATORS  # noqa: E402


def main():
    workload, spp, n_ctx = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    frames = int(sys.argv[4]) if len(sys.argv) > 4 else 3
    lib = C.CDLL(str(ROOT / "simplepath_b200" / "csrc" / "libspcu.so"))
    vp = C.c_void_p
    lib.spcu_create.argtypes = [C.c_int, C.POINTER(vp)]
    lib.spcu_upload_scene.argtypes = [vp, C.POINTER(FlatScene), C.POINTER(C.c_float), C.c_uint32]
    lib.spcu_render_frame.argtypes = [vp, C.POINTER(Partition), vp, vp, C.POINTER(Stats)]
    lib.spcu_set_option.argtypes = [vp, C.c_uint32, C.c_uint32]
    lib.spcu_last_error.argtypes = [vp]
    lib.spcu_last_error.restype = C.c_char_p
    flat = host.workload(bench.WORKLOADS[workload][0])
    jitter = np.ascontiguousarray(host.jitter(spp), dtype=np.float32)
    ctxs, rgbs, parts, stats = [], [], [], []
    for k in range(n_ctx):
        h = vp()
        assert lib.spcu_create(0, C.byref(h)) == 0
        lib.spcu_set_option(h, 3, 1)
        assert lib.spcu_upload_scene(h, flat.pointer(), jitter.ctypes.data_as(C.POINTER(C.c_float)), spp) == 0, lib.spcu_last_error(h)
        ctxs.append(h)
        rgbs.append(np.empty((flat.height, flat.width, 3), dtype=np.float32))
        lo, hi = spp * k // n_ctx, spp * (k + 1) // n_ctx
        parts.append(Partition(0, 1, lo, hi, spp, INTEGRATORS[bench.INTEGRATOR], 0))
        stats.append(Stats())

    def render(k):
        rc = lib.spcu_render_frame(ctxs[k], C.byref(parts[k]), rgbs[k].ctypes.data_as(vp), None, C.byref(stats[k]))
        assert rc == 0, lib.spcu_last_error(ctxs[k])

    wall = []
    for f in range(frames + 2):
        ts = [threading.Thread(target=render, args=(k,)) for k in range(n_ctx)]
        t0 = time.perf_counter()
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        if f >= 2:
            wall.append((time.perf_counter() - t0) * 1e3)
    paths = sum(s.paths for s in stats)
    print(json.dumps({"workload": workload, "spp": spp, "contexts": n_ctx, "grid_divisor": os.environ.get("SPCU_GRID_DIVISOR", "1"),
                      "wall_ms_min": min(wall), "wall_ms": wall, "device_ms_each": [s.device_ms for s in stats],
                      "mpaths_per_s_wall": paths / (min(wall) * 1e-3) / 1e6,
                      "mean_radiance": float(sum(r.sum() for r in rgbs) / rgbs[0].size / spp)}), flush=True)


if __name__ == "__main__":
    main()
