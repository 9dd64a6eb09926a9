#!/usr/bin/env python
"""Prints the headline figures of a bench.py JSON line: python profiles/scripts/show_bench.py FILE.json"""
import json
import sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("metric", "value", "unit", "n_gpus", "ms_per_step", "steps", "gpu_launches", "scaling", "dtype")})
print("e2e", d["e2e"]["value"], "mean radiance", d["e2e"].get("mean_radiance_of_the_frame"), "mrays/s", d.get("mrays_per_s"))
r = d.get("roofline") or {}
print("roofline", {k: r.get(k) for k in ("kernel", "bound", "frac", "issue_frac", "hbm_frac", "share_of_step", "launch_ms", "lanes_per_instruction")})
print("cpu", d.get("cpu_baseline"))
print("clocks", d.get("clocks"))
print("stages", d.get("stages_ms_per_step"))
for k, v in (d.get("configs") or {}).items():
    if "error" in v:
        print(k, v)
        continue
    rr = v["roofline"]
    print(k, round(v["value"], 1), "e2e", round(v["e2e"]["value"], 1), "ms", round(v["ms_per_step"], 2), "mrays", round(v["mrays_per_s"]),
          rr.get("kernel"), rr.get("bound"), "issue", round(rr.get("issue_frac") or 0, 3), "hbm", round(rr.get("hbm_frac") or 0, 3),
          "cpu", (v.get("cpu_baseline") or {}).get("value"))
print("construction", d.get("construction_side"))
