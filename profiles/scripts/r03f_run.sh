set -x
mkdir -p gpurun_out
CUR=simplepath_b200/csrc/libspcu.so
for w in bunny_1080p_256spp elf_1080p_256spp; do
  for cfg in 1:16777216 4:16777216 4:8388608 8:8388608 8:4194304 4:4194304; do
    SPCU_AB_LANES=${cfg%%:*} SPCU_AB_WAVEFRONT=${cfg#*:} timeout 900 python profiles/scripts/ab_frame.py $CUR $w 32 ordered 3 >> gpurun_out/r03f_ab.jsonl 2>> gpurun_out/r03f_ab.err
  done
done
tail -n 3 gpurun_out/r03f_ab.err
