"""Times spcu_build_bvh (BVHAccelerator::construct on the device) on n random boxes and, beside it, the CPU builds of the
same input on a bounded sample: the oracle's sequential restatement and — when oracle/_ref travelled — the reference's own
BVHAccelerator (single-threaded recursion over shared_ptr<Hitable>, shapes/BVHAccelerator.h:175-209).  One JSON line.

    python profiles/build_probe.py [--n 28055742] [--reps 3] [--cpu-n 2805574]
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


def boxes(n: int, seed: int = 5) -> np.ndarray:
    rng = np.random.default_rng(seed)
    c = rng.random((n, 3), dtype=np.float32) * np.float32(1000.0)
    h = rng.random((n, 3), dtype=np.float32) * np.float32(0.05)
    return np.concatenate([c - h, c + h], axis=1)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=28_055_742)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--cpu-n", type=int, default=2_805_574)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    from simplepath_b200 import capi
    ctx = capi.Context(0)
    b = boxes(args.n)
    ms, wall = [], []
    for _ in range(args.reps):
        t0 = time.perf_counter()
        got = ctx.build_bvh(b, None, 1)
        wall.append((time.perf_counter() - t0) * 1e3)
        ms.append(got["device_ms"])
    line = {"what": "spcu_build_bvh", "n": args.n, "internal_nodes": got["head"]["n_nodes"], "depth": got["head"]["max_depth"],
            "device_ms": ms, "call_ms_host_buffers": wall,
            "mprims_per_s_device": args.n / (min(ms) * 1e-3) / 1e6}
    if not args.no_cpu:
        from oracle import port, ref
        sb = boxes(args.cpu_n)
        t0 = time.perf_counter()
        want = port.build_bvh(sb, None, 1)
        line["oracle_port"] = {"n": args.cpu_n, "ms": (time.perf_counter() - t0) * 1e3}
        small = ctx.build_bvh(sb, None, 1)
        line["oracle_port"]["identical_to_device"] = bool(small["nodes"].tobytes() == want["nodes"].tobytes()
                                                          and np.array_equal(small["order"], want["order"]))
        line["device_same_n"] = {"n": args.cpu_n, "device_ms": small["device_ms"]}
        if ref.available():
            t0 = time.perf_counter()
            r = ref.build_bvh(sb, None, 1)
            line["reference"] = {"n": args.cpu_n, "ms_including_object_creation_and_flatten": (time.perf_counter() - t0) * 1e3,
                                 "identical_to_device": bool(small["nodes"].tobytes() == r["nodes"].tobytes()
                                                             and np.array_equal(small["order"], r["order"]))}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
