set -x
mkdir -p gpurun_out
for lanes in 1 0; do
  timeout 900 python bench.py --steps 3 --no-cpu --no-side-configs --batch-lanes $lanes > gpurun_out/r03o_bench_lanes$lanes.json 2> gpurun_out/r03o_bench_lanes$lanes.err
  tail -n 2 gpurun_out/r03o_bench_lanes$lanes.err
done
