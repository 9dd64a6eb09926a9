#!/usr/bin/env python
"""A/B timing of library builds on one workload, using only entry points every build since round 1 has (so an old libspcu.so
can be timed beside a new one on the same GPU in the same minute):

    python profiles/scripts/ab_frame.py LIB[,LIB...] WORKLOAD SPP [traversal: exact|ordered|default] [frames]

Per library: spcu_stats.device_ms (CUDA events around the wavefront loop, inside the library) of `frames` frames after two
warm-up frames, and the per-stage event times of one extra frame."""
import ctypes as C
import json
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from simplepath_b200 import host  # noqa: E402
from simplepath_b200.capi import FlatScene, Partition, Stats, StageTime, INTEGRATORS  # noqa: E402


def run(lib_path, flat, jitter, spp, traversal, frames):
    lib = C.CDLL(str(lib_path))
    vp = C.c_void_p
    lib.spcu_create.argtypes = [C.c_int, C.POINTER(vp)]
    lib.spcu_upload_scene.argtypes = [vp, C.POINTER(FlatScene), C.POINTER(C.c_float), C.c_uint32]
    lib.spcu_render_frame.argtypes = [vp, C.POINTER(Partition), vp, vp, C.POINTER(Stats)]
    lib.spcu_set_option.argtypes = [vp, C.c_uint32, C.c_uint32]
    lib.spcu_stage_times.argtypes = [vp, C.POINTER(StageTime), C.c_uint32, C.POINTER(C.c_uint32)]
    lib.spcu_last_error.argtypes = [vp]
    lib.spcu_last_error.restype = C.c_char_p
    lib.spcu_destroy.argtypes = [vp]
    h = vp()
    assert lib.spcu_create(0, C.byref(h)) == 0
    if traversal != "default":
        lib.spcu_set_option(h, 3, 1 if traversal == "ordered" else 0)
    if os.environ.get("SPCU_AB_WAVEFRONT"):
        lib.spcu_set_wavefront_size.argtypes = [vp, C.c_uint64]
        lib.spcu_set_wavefront_size(h, int(os.environ["SPCU_AB_WAVEFRONT"]))
    lanes = int(os.environ.get("SPCU_AB_LANES", "0"))
    if lanes:
        lib.spcu_set_option(h, 5, lanes)  # SPCU_OPT_BATCH_LANES (libraries that predate it refuse the option: ignored)
    rc = lib.spcu_upload_scene(h, flat.pointer(), jitter.ctypes.data_as(C.POINTER(C.c_float)), spp)
    assert rc == 0, lib.spcu_last_error(h)
    rgb = np.empty((flat.height, flat.width, 3), dtype=np.float32)
    part = Partition(0, 1, 0, spp, spp, INTEGRATORS[bench.INTEGRATOR], 0)
    st = Stats()
    ms = []
    for i in range(frames + 2):
        rc = lib.spcu_render_frame(h, C.byref(part), rgb.ctypes.data_as(vp), None, C.byref(st))
        assert rc == 0, lib.spcu_last_error(h)
        if i >= 2:
            ms.append(st.device_ms)
    lib.spcu_set_option(h, 1, 1)
    lib.spcu_render_frame(h, C.byref(part), rgb.ctypes.data_as(vp), None, C.byref(st))
    arr = (StageTime * 16)()
    n = C.c_uint32()
    lib.spcu_stage_times(h, arr, 16, C.byref(n))
    stages = {arr[i].name.decode(): round(arr[i].ms, 2) for i in range(n.value) if arr[i].launches}
    lib.spcu_destroy(h)
    return {"lib": Path(lib_path).name, "lanes": lanes, "wavefront": int(os.environ.get("SPCU_AB_WAVEFRONT", "0")), "device_ms_mean": float(np.mean(ms)), "device_ms_min": float(np.min(ms)),
            "mpaths_per_s": st.paths / (float(np.min(ms)) * 1e-3) / 1e6, "mean_radiance": float(rgb.mean() / spp), "stages_ms": stages}


def main():
    libs, workload, spp = sys.argv[1].split(","), sys.argv[2], int(sys.argv[3])
    traversal = sys.argv[4] if len(sys.argv) > 4 else "default"
    frames = int(sys.argv[5]) if len(sys.argv) > 5 else 3
    flat = host.workload(bench.WORKLOADS[workload][0])
    jitter = np.ascontiguousarray(host.jitter(spp), dtype=np.float32)
    for lib in libs:
        print(json.dumps({"workload": workload, "spp": spp, "traversal": traversal, **run(lib, flat, jitter, spp, traversal, frames)}),
              flush=True)


if __name__ == "__main__":
    main()
