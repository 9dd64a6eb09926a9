set -x
mkdir -p gpurun_out
V=build/variants
CUR=simplepath_b200/csrc/libspcu.so
for w in bunny_1080p_256spp elf_1080p_256spp; do
  SPCU_AB_LANES=4 timeout 900 python profiles/scripts/ab_frame.py $CUR,$V/libspcu_rf2.so,$V/libspcu_rf6.so,$V/libspcu_rf8.so,$V/libspcu_rf12.so,$CUR $w 64 ordered 2 >> gpurun_out/r04b_ab.jsonl 2>> gpurun_out/r04b_ab.err
done
tail -n 3 gpurun_out/r04b_ab.err
