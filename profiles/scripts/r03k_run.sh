set -x
mkdir -p gpurun_out
CUR=simplepath_b200/csrc/libspcu.so
for w in bunny_1080p_256spp elf_1080p_256spp; do
  SPCU_AB_LANES=4 timeout 900 python profiles/scripts/ab_frame.py $CUR $w 64 ordered 2 >> gpurun_out/r03k_ab.jsonl 2>> gpurun_out/r03k_ab.err
  SPCU_LANE_PRIORITY=1 SPCU_AB_LANES=4 timeout 900 python profiles/scripts/ab_frame.py $CUR $w 64 ordered 2 >> gpurun_out/r03k_ab_prio.jsonl 2>> gpurun_out/r03k_ab.err
done
for lanes in 4 6 8; do
  timeout 900 python bench.py --workload lucy_4k_256spp --spp 16 --steps 2 --no-cpu --no-side-configs --batch-lanes $lanes > gpurun_out/r03k_lucy_lanes$lanes.json 2> gpurun_out/r03k_lucy_lanes$lanes.err
done
tail -n 3 gpurun_out/r03k_ab.err
