set -x
mkdir -p gpurun_out
V=build/variants
CUR=simplepath_b200/csrc/libspcu.so
for w in bunny_1080p_256spp elf_1080p_256spp; do
  SPCU_AB_LANES=4 timeout 900 python profiles/scripts/ab_frame.py $V/libspcu_rf8.so,$V/libspcu_rf12.so,$V/libspcu_rf16.so,$V/libspcu_rf20.so,$V/libspcu_rf24.so,$V/libspcu_rf28.so,$V/libspcu_rf12.so $w 64 ordered 2 >> gpurun_out/r04c_ab.jsonl 2>> gpurun_out/r04c_ab.err
done
tail -n 3 gpurun_out/r04c_ab.err
