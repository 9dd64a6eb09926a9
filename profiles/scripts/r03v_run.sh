set -x
mkdir -p gpurun_out
for cfg in 0:0 8:8388608; do
  L=${cfg%%:*}; W=${cfg#*:}
  timeout 900 python bench.py --steps 5 --no-cpu --no-side-configs --batch-lanes $L --wavefront-size $W > gpurun_out/r03v_bunny_${L}_$W.json 2>> gpurun_out/r03v.err
  timeout 900 python bench.py --workload elf_1080p_256spp --steps 3 --no-cpu --no-side-configs --batch-lanes $L --wavefront-size $W > gpurun_out/r03v_elf_${L}_$W.json 2>> gpurun_out/r03v.err
  timeout 900 python bench.py --workload lucy_4k_256spp --spp 16 --steps 2 --no-cpu --no-side-configs --batch-lanes $L --wavefront-size $W > gpurun_out/r03v_lucy_${L}_$W.json 2>> gpurun_out/r03v.err
done
tail -n 3 gpurun_out/r03v.err
