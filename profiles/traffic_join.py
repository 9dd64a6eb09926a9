#!/usr/bin/env python
"""ncu launch list (per-launch metrics, CSV) + traffic_probe.py's item counts -> per-item figures per wavefront stage:
DRAM bytes, thread instructions, warp instructions (-> active lanes per instruction) and the time-weighted issue-slot
utilisation.  bench.py reads the result (profiles/ncu_kernels_rNN_<config>.json) for `roofline`.

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__thread_inst_executed.sum,\
smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file launches.csv \
        python profiles/traffic_probe.py WORKLOAD SPP > probe.json
    python profiles/traffic_join.py launches.csv probe.json out.json
"""
import csv
import json
import re
import sys

KERNEL_TO_STAGE = {"k_raygen": "raygen", "k_extend": "extend", "k_shade": "shade", "k_nee_light": "nee_light",
                   "k_shadow": "shadow", "k_nee_bsdf": "nee_bsdf", "k_mis_trace": "mis_trace",
                   "k_nee_mis_accumulate": "nee_mis_accumulate", "k_direct_accumulate": "direct_accumulate",
                   "k_advance": "advance", "k_resolve": "resolve", "k_paths": "paths", "k_smwave": "paths",
                   "k_extend_begin": "extend", "k_extend_walk": "extend", "k_shadow_begin": "shadow", "k_shadow_walk": "shadow"}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}


def main() -> None:
    launches, probe, out = sys.argv[1:4]
    info = json.loads(open(probe).read().strip().splitlines()[-1])
    items = {s["name"]: s["items"] for s in info["stages"]}
    with open(launches) as f:
        rows = list(csv.reader(l for l in f if l.startswith('"')))
    idx = {h: i for i, h in enumerate(rows[0])}
    agg: dict[str, dict[str, float]] = {}
    for r in rows[1:]:
        m = re.search(r"\b(k_[a-z_0-9]+)", r[idx["Kernel Name"]])   # "void spcu::<unnamed>::k_extend_walk<0, 1, spcu::FeatFull>(...)"
        stage = KERNEL_TO_STAGE.get(m.group(1)) if m else None
        if stage is None:
            continue
        a = agg.setdefault(stage, {"dram_bytes": 0.0, "seconds": 0.0, "launches": 0, "thread_inst": 0.0, "warp_inst": 0.0,
                                   "issue_pct_x_s": 0.0, "by_id": {}})
        v = float(r[idx["Metric Value"]].replace(",", "")) * SCALE.get(r[idx["Metric Unit"]], 1.0)
        metric = r[idx["Metric Name"]]
        launch = a["by_id"].setdefault(r[idx["ID"]], {})
        launch[metric] = v
        if metric.startswith("dram__bytes"):
            a["dram_bytes"] += v
        elif metric.startswith("gpu__time_duration"):
            a["seconds"] += v
            a["launches"] += 1
        elif metric.startswith("smsp__thread_inst_executed"):
            a["thread_inst"] += v
        elif metric.startswith("smsp__inst_executed"):
            a["warp_inst"] += v
    for a in agg.values():   # issue-slot utilisation: weighted by each launch's duration
        for launch in a.pop("by_id").values():
            pct = launch.get("smsp__issue_active.avg.pct_of_peak_sustained_active")
            if pct is not None:
                a["issue_pct_x_s"] += pct * launch.get("gpu__time_duration.sum", 0.0)
    result = {"capture": launches, "probe": {k: info[k] for k in ("workload", "spp", "pipeline", "traversal")}, "stages": {}}
    for stage, a in agg.items():
        n = items.get(stage, 0)
        result["stages"][stage] = {"dram_bytes_per_item": a["dram_bytes"] / n if n else None, "items": n,
                                   "launches": a["launches"], "dram_bytes": a["dram_bytes"],
                                   "ncu_seconds": a["seconds"],
                                   "thread_inst_per_item": a["thread_inst"] / n if n else None,
                                   "warp_inst_per_item": a["warp_inst"] / n if n else None,
                                   "lanes_per_instruction": a["thread_inst"] / a["warp_inst"] if a["warp_inst"] else None,
                                   "issue_active_pct": a["issue_pct_x_s"] / a["seconds"] if a["seconds"] and a["issue_pct_x_s"] else None}
    json.dump(result, open(out, "w"), indent=1)
    for stage, v in result["stages"].items():
        print(f"{stage:20s} {v['dram_bytes_per_item'] or 0:8.1f} B/item {v['thread_inst_per_item'] or 0:9.0f} inst/item "
              f"{v['lanes_per_instruction'] or 0:5.1f} lanes {v['issue_active_pct'] or 0:5.1f} % issue  {v['items']:>12d} items "
              f"{v['launches']:4d} launches {v['ncu_seconds'] * 1e3:8.2f} ms")


if __name__ == "__main__":
    main()
