"""Times the construction side end to end on the device at lucy's size: spcu_ingest_mesh (face / vertex normals, Mesh
pre-transform, triangle records) and spcu_upload_scene_build (bounds, BVH, leaf-order gather) on a procedural mesh of n
triangles, with the oracle's CPU restatement of both on a bounded sample beside it.  One JSON line.

    python profiles/ingest_probe.py [--n 28055742] [--cpu-n 2805574]
"""
from __future__ import annotations

import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=28_055_742)
    ap.add_argument("--cpu-n", type=int, default=2_805_574)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    from simplepath_b200 import capi, scenes
    from simplepath_b200.flat import FlatSceneData
    import meshcases
    ctx = capi.Context(0)
    xf = meshcases.transform()
    m = xf[:9].reshape(3, 3).T.astype(np.float64)
    nxf = np.linalg.inv(m).reshape(9).astype(np.float32)

    def run(n, who):
        v, f = scenes.bumpy_sphere(n, (-0.1, 0.03, -0.06), (0.06, 0.19, 0.06))
        v, f = np.asarray(v, dtype=np.float32), np.asarray(f, dtype=np.uint32)
        t0 = time.perf_counter()
        r = who(v, f)
        return v, f, r, (time.perf_counter() - t0) * 1e3

    v, f, r, call_ms = run(args.n, lambda v, f: ctx.ingest_mesh(v, f, xf, nxf, 0))
    line = {"what": "spcu_ingest_mesh + spcu_upload_scene_build", "faces": int(len(f)), "vertices": int(len(v)),
            "triangles_kept": int(len(r["prims"])), "ingest_device_ms": r["device_ms"], "ingest_call_ms_host_buffers": call_ms}
    # a scene shell (camera, plane, lights, materials) around the mesh
    shell = FlatSceneData.load(ROOT / "tests" / "golden" / "g_bunny.flat.npz")
    nu = shell.head["geom"]["n_unbounded"]
    k = len(r["prims"])
    for name, rec, width in (("geom_prims", r["prims"], 48), ("geom_shade", r["shade"], 48), ("geom_meta", r["meta"], 4)):
        shell.arrays[name] = np.concatenate([shell.arrays[name][:nu], np.ascontiguousarray(rec).view(np.uint8).reshape(k, width)])
    shell.head["geom"] = dict(shell.head["geom"], n_prims=nu + k, n_nodes=0, root=~nu, root_count=0, max_depth=0)
    shell.arrays["geom_nodes"] = np.zeros((0, 64), dtype=np.uint8)
    jitter = np.zeros((1, 2), dtype=np.float32)
    t0 = time.perf_counter()
    order, head = ctx.upload_scene_build(shell.pointer(), jitter, keepalive=shell)
    line["upload_scene_build_call_ms"] = (time.perf_counter() - t0) * 1e3
    line["accel"] = head
    line["scene_bytes"] = ctx.scene_bytes()
    if not args.no_cpu:
        from oracle import port
        sv, sf, want, ms = run(args.cpu_n, lambda v, f: port.ingest_mesh(v, f, xf, nxf, 0))
        got = ctx.ingest_mesh(sv, sf, xf, nxf, 0)
        line["oracle_ingest"] = {"faces": int(len(sf)), "ms": ms, "device_ms_same_input": got["device_ms"],
                                 "positions_identical": bool(got["prims"].tobytes() == want["prims"].tobytes())}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
