#!/usr/bin/env python
"""Prints an ab_frame.py record file as a table: python profiles/scripts/show_ab.py FILE.jsonl"""
import json
import sys
for line in open(sys.argv[1]):
    d = json.loads(line)
    st = d["stages_ms"]
    print(f"{d['workload'][:6]:6s} spp {d['spp']:3d} lanes {d.get('lanes', 0)} {d['lib']:24s} {d['device_ms_min']:8.2f} ms {d['mpaths_per_s']:7.1f} Mp/s  rad {d['mean_radiance']:.8f}  " +
          " ".join(f"{k}={st[k]}" for k in ("extend", "shade", "shadow", "nee_bsdf", "mis_trace", "advance") if k in st))
