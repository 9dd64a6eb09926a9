#!/usr/bin/env python
"""Benchmark of the hot path: whole-frame renders of the BASELINE.json workload through the CUDA backend.

    python bench.py --gpus N --steps K --warmup W            # this repo's CudaIntegrator path (one rank per GPU)
    python bench.py --impl reference --steps K --warmup W    # the reference's own CPU implementation, same workload

A *step* is one frame: `spp` samples for every pixel of the workload's image on every rank.  At N = 1 the workload is
BASELINE.json configs[1] (example_scene, 1920x1080, 64 spp).  With N ranks the sample range is partitioned (rank r
renders global samples [r*spp, (r+1)*spp) of N*spp, scene replicated, RNG keyed by the global sample index) and the
per-pixel accumulators are summed to rank 0 with NCCL inside the timed region: per-GPU work is fixed ("weak").

value   = Mpaths/s, device-timed (CUDA events, max over ranks), scene resident in HBM, accumulators on the device
e2e     = the same metric through the host-buffer C-ABI call a plugin makes (spcu_upload_scene + spcu_render):
          flattened scene copied host->device and accumulators copied device->host inside the timed region
roofline= the kernel with the largest share of the timed region, algorithmic bytes (DESIGN.md "byte model") over its
          mean CUDA-event duration, against MEASURED_PEAKS.json
cpu_baseline = the reference binary (oracle/_ref/SimplePath) on this box's host cores on a bounded sample of the workload
"""
from __future__ import annotations

import argparse
import json
import os
import re
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    # name -> (scene name in simplepath_b200.scenes, spp)
    "example_scene_1080p_64spp": ("c2_example_scene", 64),
    "material_spheres_256_16spp": ("c1_material_spheres_const", 16),
    "bunny_1080p_256spp": ("c3_bunny", 256),
    "elf_1080p_256spp": ("c4_elf", 256),
    "lucy_4k_256spp": ("c5_lucy", 256),
}
DEFAULT_WORKLOAD = "example_scene_1080p_64spp"
INTEGRATOR = "iterative_rrnee"

# ---- byte model: algorithmic bytes per item of each wavefront stage, excluding traversal ---------------
STAGE_BYTES = {
    # 32-byte records of csrc/device_scene.h: what a stage must read and write per queue entry
    "raygen": 88,               # pixel id in; RayRec, PathRec, RNG counter, queue entry out
    "extend": 72,               # queue, RayRec in; ExtendRec, sorted-queue entry out (+ traversal bytes)
    "shade": 240,               # queue, ExtendRec, RayRec, PathRec, shading record + meta in; VertexRec, SampleRec, queue out
    "nee_light": 108,           # queue, VertexRec, PathRec in; LightRec, counter, queue out
    "shadow": 56,               # queue, VertexRec.p, LightRec in; queue out (+ traversal bytes)
    "nee_bsdf": 188,            # queue, VertexRec, LightRec, PathRec, RayRec.d in; MisRec, counter, queue out
    "mis_trace": 44,            # queue, VertexRec.p, MisRec.d in; light/occluded out (+ traversal bytes)
    "nee_mis_accumulate": 116,  # queue, MisRec, PathRec in; PathRec.L out
    "direct_accumulate": 116,
    "advance": 156,             # queue, SampleRec, VertexRec, PathRec in; PathRec.tp, RayRec, counter, queue out
    "resolve": 20,              # radiance sample in, per-pixel sums amortised
    "paths": 20,  # persistent path kernel: 4 B pixel id in, 16 B radiance sample out; everything else stays on chip
}
NODE_BYTES, TRI_BYTES, XF_BYTES = 64, 48, 96


def measured_traffic_per_item(stage: str, workload: str, pipeline: str | None = None) -> tuple[float | None, str | None]:
    """DRAM bytes per item of a stage as ncu measured them (dram__bytes_read.sum + dram__bytes_write.sum over every
    launch of one render of this workload, divided by the stage's items: profiles/traffic_probe.py + traffic_join.py)."""
    for path in sorted((ROOT / "profiles").glob("ncu_traffic_r*.json"), reverse=True):
        d = json.loads(path.read_text())
        if pipeline is not None and d.get("probe", {}).get("pipeline") != pipeline:
            continue
        if d.get("probe", {}).get("workload") == workload and stage in d.get("stages", {}):
            v = d["stages"][stage].get("dram_bytes_per_item")
            if v is not None:
                return float(v), path.name
    return None, None


def peaks() -> tuple[float, str]:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def construction_side(ctx) -> dict:
    """Outside the timed region and not part of `value`: the steps either side of the path on a bounded
    synthetic input, device-timed — BVHAccelerator::construct (spcu_build_bvh), read_ply's normal passes + Mesh's constructor
    (spcu_ingest_mesh) — and the same inputs through the oracle's sequential restatement on one host core."""
    from oracle import port
    from simplepath_b200 import scenes
    n = 1_000_000
    rng = np.random.default_rng(5)
    c = rng.random((n, 3), dtype=np.float32) * np.float32(100.0)
    h = rng.random((n, 3), dtype=np.float32) * np.float32(0.05)
    boxes = np.concatenate([c - h, c + h], axis=1)
    ctx.build_bvh(boxes)  # warm-up (allocations, kernel load)
    built = ctx.build_bvh(boxes)
    t0 = time.perf_counter()
    want = port.build_bvh(boxes)
    cpu_build_ms = (time.perf_counter() - t0) * 1e3
    v, f = scenes.bumpy_sphere(n, (-0.1, 0.03, -0.06), (0.06, 0.19, 0.06))
    v, f = np.asarray(v, dtype=np.float32), np.asarray(f, dtype=np.uint32)
    xf = np.array([10, 0, 0, 0, 10, 0, 0, 0, 10, 0, 0, 0], dtype=np.float32)
    nxf = np.array([0.1, 0, 0, 0, 0.1, 0, 0, 0, 0.1], dtype=np.float32)
    ctx.ingest_mesh(v, f, xf, nxf)
    ingested = ctx.ingest_mesh(v, f, xf, nxf)
    t0 = time.perf_counter()
    want_mesh = port.ingest_mesh(v, f, xf, nxf)
    cpu_ingest_ms = (time.perf_counter() - t0) * 1e3
    return {
        "bvh_build": {"primitives": n, "device_ms": built["device_ms"], "internal_nodes": built["head"]["n_nodes"],
                      "oracle_port_ms_1_core": cpu_build_ms,
                      "identical_to_oracle": bool(built["nodes"].tobytes() == want["nodes"].tobytes()
                                                  and np.array_equal(built["order"], want["order"]))},
        "mesh_ingest": {"faces": int(len(f)), "vertices": int(len(v)), "device_ms": ingested["device_ms"],
                        "oracle_port_ms_1_core": cpu_ingest_ms,
                        "positions_identical_to_oracle": bool(ingested["prims"].tobytes() == want_mesh["prims"].tobytes())},
    }


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# =====================================================================================================================
def run_cuda(args) -> None:
    import torch
    import torch.distributed as dist
    from simplepath_b200 import capi, distributed, host
    if not capi.LIB_PATH.exists() and int(os.environ.get("LOCAL_RANK", "0")) == 0:
        import __graft_entry__   # the library normally travels with the snapshot; build it here if it did not
        __graft_entry__.build()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    scene_name, spp = WORKLOADS[args.workload]
    spp = args.spp or spp
    flat = host.workload(scene_name)
    w, h = flat.width, flat.height
    spp_total = spp * world
    jitter = host.jitter(spp_total)
    ctx = capi.Context(local_rank)
    ctx.set_option(capi.OPT_PIPELINE, {"auto": capi.PIPELINE_AUTO, "smwave": capi.PIPELINE_SMWAVE, "paths": capi.PIPELINE_PATHS,
                                       "wavefront": capi.PIPELINE_WAVEFRONT}[args.pipeline])
    ctx.set_option(capi.OPT_TRAVERSAL, capi.TRAVERSAL_ORDERED if args.traversal == "ordered" else capi.TRAVERSAL_EXACT)
    ctx.upload_scene(flat.pointer(), jitter, keepalive=flat)
    part = distributed.sample_partition(rank, world, spp, INTEGRATOR, args.seed)

    rgb = torch.zeros((h, w, 3), dtype=torch.float32, device="cuda")
    sq = torch.zeros((h, w), dtype=torch.float32, device="cuda")
    stream = torch.cuda.current_stream()

    def step(want_stats: bool):
        rgb.zero_(); sq.zero_()
        st = ctx.render_device(part, rgb.data_ptr(), sq.data_ptr(), stream.cuda_stream, want_stats=want_stats)
        distributed.reduce_to_root(rgb, sq)   # NCCL over NVLink, inside the timed region
        return st

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (also: one counting pass for the byte model's N_node / N_tri / N_xf) --------------------------------
    ctx.set_option(capi.OPT_COUNT_NODES, 1)
    counted = step(True)
    ctx.set_option(capi.OPT_COUNT_NODES, 0)
    for _ in range(max(args.warmup - 1, 2)):
        step(False)
    barrier()

    # ---- timed region: K steps, device-timed ----------------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step(False)          # nothing but kernel launches (and the NCCL reduce) between the two events
    e1.record(stream)
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())

    # ---- per-kernel CUDA-event durations, same workload, directly after the timed steps (one event pair per launch
    # would perturb the headline number in the wavefront pipeline, so they are taken on their own steps) -----------
    ctx.set_option(capi.OPT_STAGE_TIMING, 1)
    stage_ms: dict[str, float] = {}
    stage_launches: dict[str, int] = {}
    stage_items: dict[str, int] = {}
    last = None
    stage_steps = max(1, min(args.steps, 3))
    for _ in range(stage_steps):
        last = step(True)
        for s in ctx.stage_times():
            stage_ms[s["name"]] = stage_ms.get(s["name"], 0.0) + s["ms"]
            stage_launches[s["name"]] = stage_launches.get(s["name"], 0) + s["launches"]
            stage_items[s["name"]] = stage_items.get(s["name"], 0) + s["items"]
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ctx.set_option(capi.OPT_STAGE_TIMING, 0)

    counts = torch.tensor([last["paths"], last["rays_closest"], last["rays_any"], last["rays_lights"],
                           last["kernel_launches"]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(counts)
    paths, rays_closest, rays_any, rays_lights, launches = (float(x) for x in counts.tolist())

    # ---- end to end through the host-buffer C-ABI (what sp::CudaIntegrator::render_frame does) -------------------
    # page-locked result buffers (the device->host read of the step's result runs at PCIe speed)
    h_rgb = torch.empty((h, w, 3), dtype=torch.float32, pin_memory=True).numpy()
    h_sq = torch.empty((h, w), dtype=torch.float32, pin_memory=True).numpy()
    e2e_steps = max(1, min(args.steps, 5))

    def e2e_step():
        ctx.upload_scene(flat.pointer(), jitter, keepalive=flat)   # host -> device: the step's inputs (scene + jitter table)
        ctx.render_frame(part, out=(h_rgb, h_sq))                  # device -> host: the step's result (overwritten, not accumulated)

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_s.item())
    scene_bytes = ctx.scene_bytes()
    acc_bytes = h_rgb.nbytes + h_sq.nbytes

    if rank != 0:
        ctx.close()
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ------------------------------------------------------------------------------
    peak, peak_src = peaks()
    dominant = max(stage_ms, key=stage_ms.get)
    n_launch = max(stage_launches[dominant], 1)
    items = stage_items[dominant]
    bytes_total = items * STAGE_BYTES[dominant]
    trav = {"extend", "shadow", "mis_trace", "paths"}
    geom_queries = counted["rays_closest"] + counted["rays_any"]
    per_query = ((counted["nodes_visited"] * NODE_BYTES + counted["prims_tested"] * TRI_BYTES +
                  counted["xf_prims_tested"] * XF_BYTES) / max(geom_queries, 1))
    if dominant == "paths":  # one item = one path = (closest + any-hit) geometry queries
        bytes_total += geom_queries * per_query * (items / max(counted["paths"], 1))
    elif dominant in trav:
        bytes_total += items * per_query
    dur_s = stage_ms[dominant] / 1e3 / n_launch
    achieved = bytes_total / n_launch / dur_s / 1e9 if dur_s > 0 else 0.0
    kernel_ms = sum(stage_ms.values())
    traffic_item, traffic_src = measured_traffic_per_item(dominant, args.workload, ctx.resolved_pipeline())

    cpu = cpu_baseline(scene_name, flat) if world == 1 and not args.no_cpu else None
    construction = construction_side(ctx) if world == 1 and not args.no_cpu else None

    steps_s = total_ms / 1e3
    line = {
        "metric": "Mpaths/s", "value": paths * args.steps / steps_s / 1e6, "unit": "Mpaths/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "scene": scene_name, "width": w, "height": h, "spp_per_gpu": spp,
                   "integrator": INTEGRATOR, "pipeline": ctx.resolved_pipeline(), "pipeline_option": args.pipeline, "traversal": args.traversal, "max_depth": flat.head["max_depth"], "rr_depth": flat.head["rr_depth"],
                   "paths_per_step": int(paths), "partition": f"sample ranges x{world}, scene replicated",
                   "l2": "no explicit flush: each step streams >1 GB of wavefront state, far above the 126 MB L2"},
        "mrays_per_s": (rays_closest + rays_any) * args.steps / steps_s / 1e6,
        "rays": {"closest_per_path": rays_closest / paths, "any_hit_per_path": rays_any / paths,
                 "lights_accel_per_path": rays_lights / paths},
        "e2e": {"value": paths * e2e_steps / e2e_s / 1e6, "unit": "Mpaths/s", "h2d_bytes_per_step": scene_bytes,
                "d2h_bytes_per_step": acc_bytes, "steps": e2e_steps,
                "call": "spcu_upload_scene + spcu_render_frame (host buffers, page-locked result)"},
        "gpu_launches": int(launches / world * args.steps),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak,
                     "traffic": traffic_item * items / n_launch if traffic_item is not None else None,
                     "traffic_source": traffic_src, "peak_source": peak_src,
                     "share_of_step": stage_ms[dominant] / kernel_ms if kernel_ms else None,
                     "launches": n_launch, "items_per_launch": items / n_launch,
                     "bytes_per_item": bytes_total / max(items, 1),
                     "note": "shading stages are instruction-issue bound, not HBM bound (profiles/)"},
        "stages_ms_per_step": {k: v / stage_steps for k, v in stage_ms.items() if stage_launches.get(k)},
        "cpu_baseline": cpu,
        "construction_side": construction,
    }
    emit(line)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


# =====================================================================================================================
def ref_binary() -> Path | None:
    p = ROOT / "oracle" / "_ref" / "SimplePath"
    return p if p.exists() else None


def run_reference_binary(scene_path: Path, spp: int, threads: int, timeout: float) -> float:
    """Stock reference executable; render seconds from its own Stopwatch line (main.cpp:138-141).  The process hangs at
    exit (AccumulatedLogger), so it is killed as soon as the line is seen."""
    proc = subprocess.Popen(["stdbuf", "-o0", str(ref_binary()), "--threads", str(threads), "--samples", str(spp),
                             "--integrator", INTEGRATOR, scene_path.name],
                            cwd=scene_path.parent, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    deadline = time.time() + timeout
    secs, buf = None, ""
    try:
        while time.time() < deadline:
            ch = proc.stdout.read(1)
            if not ch:
                break
            buf += ch
            m = re.search(r"Elapsed time:\s*([^\n]+)\n", buf)
            if m:
                secs = parse_elapsed(m.group(1))
                break
    finally:
        proc.kill()
        proc.wait()
    if secs is None:
        raise RuntimeError(f"reference did not finish within {timeout}s: {buf[-300:]}")
    return secs


def parse_elapsed(text: str) -> float:
    """Stopwatch::print (base/Stopwatch.h:47-60) writes HH:MM:SS.cc (centiseconds)."""
    m = re.fullmatch(r"\s*(\d+):(\d+):(\d+)\.(\d+)\s*", text)
    if not m:
        raise ValueError(f"cannot parse elapsed time {text!r}")
    hh, mm, ss, cc = (int(g) for g in m.groups())
    return hh * 3600.0 + mm * 60.0 + ss + cc / 100.0


def reference_sample(scene_name: str, budget_s: float = 15.0):
    """Bounded sample of the workload for the CPU: same scene and resolution, reduced spp (throughput is spp
    independent), sized from a 1 spp probe to about `budget_s` seconds."""
    from simplepath_b200 import scenes
    path = scenes.ensure(scene_name)
    w, h, _ = scenes.info(scene_name)
    threads = os.cpu_count() or 1
    probe = run_reference_binary(path, 1, threads, 600)
    spp = int(max(1, min(64, budget_s / max(probe, 1e-3))))
    return path, w, h, spp, threads


def cpu_baseline(scene_name: str, flat) -> dict:
    """Reference CPU path on this box's host cores, bounded sample.  kind 'reference' = the compiled reference binary;
    'port' = the oracle's C restatement when that binary is not present."""
    threads = os.cpu_count() or 1
    if ref_binary() is not None:
        path, w, h, spp, threads = reference_sample(scene_name)
        secs = run_reference_binary(path, spp, threads, 900)
        return {"value": w * h * spp / secs / 1e6, "unit": "Mpaths/s", "cores": threads, "kind": "reference",
                "sample": f"{scene_name} {w}x{h} at {spp} spp ({w * h * spp} paths, {secs:.2f} s), stock binary "
                          f"--threads {threads} --integrator {INTEGRATOR}, its own Stopwatch"}
    from oracle import port
    from simplepath_b200 import host
    from simplepath_b200.capi import INTEGRATORS, Partition
    spp = 1
    jitter = host.jitter(spp)
    t0 = time.perf_counter()
    _, _, st = port.render(flat.pointer(), jitter, Partition(0, 1, 0, spp, spp, INTEGRATORS[INTEGRATOR], 0), threads=threads)
    secs = time.perf_counter() - t0
    return {"value": st["paths"] / secs / 1e6, "unit": "Mpaths/s", "cores": threads, "kind": "port",
            "sample": f"{scene_name} at {spp} spp ({st['paths']} paths, {secs:.2f} s), oracle C restatement, OpenMP"}


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from simplepath_b200 import scenes
    scene_name, spp_full = WORKLOADS[args.workload]
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    w, h, _ = scenes.info(scene_name)
    if ref_binary() is not None:
        path, w, h, spp, threads = reference_sample(scene_name, budget_s=10.0)
        kind = "reference"

        def one():
            return run_reference_binary(path, spp, threads, 900)
    else:
        from oracle import port
        from simplepath_b200 import host
        from simplepath_b200.capi import INTEGRATORS, Partition
        flat = host.workload(scene_name)
        spp, threads, kind = 1, os.cpu_count() or 1, "port"
        jitter = host.jitter(spp)

        def one():
            t0 = time.perf_counter()
            port.render(flat.pointer(), jitter, Partition(0, 1, 0, spp, spp, INTEGRATORS[INTEGRATOR], 0), threads=threads)
            return time.perf_counter() - t0
    for _ in range(min(args.warmup, 1)):
        one()
    steps = max(1, min(args.steps, 3))
    secs = [one() for _ in range(steps)]
    paths = w * h * spp
    value = paths * steps / sum(secs) / 1e6
    sample = (f"{scene_name} {w}x{h} at {spp} spp per step ({paths} paths/step; full workload is {spp_full} spp), "
              f"{'stock reference binary' if kind == 'reference' else 'oracle C restatement'}, {threads} threads")
    emit({
        "impl": "reference", "metric": "Mpaths/s", "value": value, "unit": "Mpaths/s", "n_gpus": world, "steps": steps,
        "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * sum(secs) / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "scene": scene_name, "width": w, "height": h, "spp_per_step": spp,
                   "integrator": INTEGRATOR},
        "cpu_baseline": {"value": value, "unit": "Mpaths/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


_REAL_STDOUT = None


def claim_stdout() -> None:
    """stdout carries exactly ONE JSON line: everything libraries print (NCCL's version banner, torchrun notices)
    goes to stderr instead; emit() writes the line to the real stdout."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line: dict) -> None:
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main() -> None:
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--spp", type=int, default=0, help="override samples per pixel per GPU")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--traversal", default="exact", choices=["exact", "ordered"],
                    help="closest-hit walk of the extend stage (SPCU_OPT_TRAVERSAL)")
    ap.add_argument("--pipeline", default="auto", choices=["auto", "smwave", "paths", "wavefront"],
                    help="kernel organisation (SPCU_OPT_PIPELINE); same estimator and random numbers either way")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "cuda":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
