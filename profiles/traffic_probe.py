#!/usr/bin/env python
"""One render of a bench workload through the C-ABI, printing the per-stage item counts as JSON.  Run it under
`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv` to get the DRAM traffic of every
launch; profiles/traffic_join.py divides the per-stage sums by these item counts (-> profiles/ncu_traffic_rNN.json, which
bench.py uses for `roofline.traffic`).

    python profiles/traffic_probe.py [workload] [spp] [pipeline] [traversal]
"""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
from simplepath_b200 import capi  # noqa: E402


def main() -> None:
    workload = sys.argv[1] if len(sys.argv) > 1 else bench.DEFAULT_WORKLOAD
    scene, spp = bench.WORKLOADS[workload]
    spp = int(sys.argv[2]) if len(sys.argv) > 2 else spp
    pipeline = sys.argv[3] if len(sys.argv) > 3 else "auto"
    traversal = sys.argv[4] if len(sys.argv) > 4 else "ordered"
    ctx = capi.Context(0)
    ctx.set_option(capi.OPT_PIPELINE, {"auto": capi.PIPELINE_AUTO, "smwave": capi.PIPELINE_SMWAVE, "paths": capi.PIPELINE_PATHS,
                                       "wavefront": capi.PIPELINE_WAVEFRONT}[pipeline])
    ctx.set_option(capi.OPT_TRAVERSAL, capi.TRAVERSAL_ORDERED if traversal == "ordered" else capi.TRAVERSAL_EXACT)
    flat, jitter, _ = bench.load_scene(ctx, scene, spp)   # (c5_lucy: generated, ingested and built on the device)
    part = capi.Partition(0, 1, 0, spp, spp, capi.INTEGRATORS[bench.INTEGRATOR], 0)
    _, _, stats = ctx.render_frame(part, want_sumsq=False)
    pipeline = ctx.resolved_pipeline()
    print(json.dumps({"workload": workload, "scene": scene, "spp": spp, "pipeline": pipeline, "traversal": traversal,
                      "stats": stats, "stages": ctx.stage_times()}))
    ctx.close()


if __name__ == "__main__":
    main()
