set -x
mkdir -p gpurun_out
for lanes in 1 0; do
  ( time timeout 900 python bench.py --workload lucy_4k_256spp --spp 8 --steps 2 --no-cpu --no-side-configs --batch-lanes $lanes > gpurun_out/r03d_lucy_lanes$lanes.json 2> gpurun_out/r03d_lucy_lanes$lanes.err ) 2> gpurun_out/r03d_lucy_lanes$lanes.time
  tail -n 2 gpurun_out/r03d_lucy_lanes$lanes.err
  nvidia-smi --query-gpu=memory.used --format=csv
done
