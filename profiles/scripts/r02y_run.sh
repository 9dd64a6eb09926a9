set -x
mkdir -p gpurun_out
P=profiles/scripts/concurrent_probe.py
for w in bunny_1080p_256spp elf_1080p_256spp; do
  timeout 600 python $P $w 24 3 >> gpurun_out/r02y_concurrent.jsonl 2>> gpurun_out/r02y.err
  timeout 600 python $P $w 32 4 >> gpurun_out/r02y_concurrent.jsonl 2>> gpurun_out/r02y.err
  timeout 600 python $P $w 32 2 >> gpurun_out/r02y_concurrent.jsonl 2>> gpurun_out/r02y.err
  timeout 600 python $P $w 8 2 >> gpurun_out/r02y_concurrent.jsonl 2>> gpurun_out/r02y.err
  timeout 600 python $P $w 8 4 >> gpurun_out/r02y_concurrent.jsonl 2>> gpurun_out/r02y.err
done
tail -n 5 gpurun_out/r02y.err
