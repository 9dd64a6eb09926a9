#!/usr/bin/env python
"""ncu launch list (dram bytes per launch, CSV) + traffic_probe.py's item counts -> DRAM bytes per item per stage.

    python profiles/traffic_join.py launches.csv probe.json out.json
"""
import csv
import json
import re
import sys

KERNEL_TO_STAGE = {"k_raygen": "raygen", "k_extend": "extend", "k_shade": "shade", "k_nee_light": "nee_light",
                   "k_shadow": "shadow", "k_nee_bsdf": "nee_bsdf", "k_mis_trace": "mis_trace",
                   "k_nee_mis_accumulate": "nee_mis_accumulate", "k_direct_accumulate": "direct_accumulate",
                   "k_advance": "advance", "k_resolve": "resolve", "k_paths": "paths"}
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
        name = re.sub(r"<.*", "", re.sub(r".*::", "", re.sub(r"\(.*", "", r[idx["Kernel Name"]])))
        stage = KERNEL_TO_STAGE.get(name)
        if stage is None:
            continue
        a = agg.setdefault(stage, {"dram_bytes": 0.0, "seconds": 0.0, "launches": 0})
        v = float(r[idx["Metric Value"]].replace(",", "")) * SCALE.get(r[idx["Metric Unit"]], 1.0)
        metric = r[idx["Metric Name"]]
        if metric.startswith("dram__bytes"):
            a["dram_bytes"] += v
        elif metric.startswith("gpu__time_duration"):
            a["seconds"] += v
            a["launches"] += 1
    result = {"capture": launches, "probe": {k: info[k] for k in ("workload", "spp", "pipeline", "traversal")}, "stages": {}}
    for stage, a in agg.items():
        n = items.get(stage, 0)
        result["stages"][stage] = {"dram_bytes_per_item": a["dram_bytes"] / n if n else None, "items": n,
                                   "launches": a["launches"], "dram_bytes": a["dram_bytes"],
                                   "ncu_seconds": a["seconds"]}
    json.dump(result, open(out, "w"), indent=1)
    for stage, v in result["stages"].items():
        print(f"{stage:20s} {v['dram_bytes_per_item'] or 0:8.1f} B/item  {v['items']:>12d} items {v['launches']:4d} launches")


if __name__ == "__main__":
    main()
