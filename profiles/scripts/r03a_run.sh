set -x
mkdir -p gpurun_out
CUR=simplepath_b200/csrc/libspcu.so
for w in bunny_1080p_256spp elf_1080p_256spp; do
  for cfg in 4:16777216 4:33554432 2:33554432 3:33554432 8:8388608 4:8388608; do
    SPCU_AB_LANES=${cfg%%:*} SPCU_AB_WAVEFRONT=${cfg#*:} timeout 900 python profiles/scripts/ab_frame.py $CUR $w 128 ordered 2 >> gpurun_out/r03a_ab.jsonl 2>> gpurun_out/r03a_ab.err
  done
done
tail -n 3 gpurun_out/r03a_ab.err
