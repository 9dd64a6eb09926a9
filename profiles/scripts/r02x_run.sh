set -x
mkdir -p gpurun_out
P=profiles/scripts/concurrent_probe.py
for w in bunny_1080p_256spp elf_1080p_256spp; do
  timeout 600 python $P $w 16 1 >> gpurun_out/r02x_concurrent.jsonl 2>> gpurun_out/r02x.err
  timeout 600 python $P $w 16 2 >> gpurun_out/r02x_concurrent.jsonl 2>> gpurun_out/r02x.err
  SPCU_GRID_DIVISOR=2 timeout 600 python $P $w 16 2 >> gpurun_out/r02x_concurrent.jsonl 2>> gpurun_out/r02x.err
  SPCU_GRID_DIVISOR=2 timeout 600 python $P $w 16 1 >> gpurun_out/r02x_concurrent.jsonl 2>> gpurun_out/r02x.err
  SPCU_GRID_DIVISOR=3 timeout 600 python $P $w 18 3 >> gpurun_out/r02x_concurrent.jsonl 2>> gpurun_out/r02x.err
done
tail -n 5 gpurun_out/r02x.err
cat gpurun_out/r02x_concurrent.jsonl
