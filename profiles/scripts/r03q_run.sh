set -x
mkdir -p gpurun_out
V=build/variants
CUR=simplepath_b200/csrc/libspcu.so
for w in bunny_1080p_256spp elf_1080p_256spp; do
  SPCU_AB_LANES=1 timeout 900 python profiles/scripts/ab_frame.py $V/libspcu_noilp.so,$CUR,$V/libspcu_noilp.so,$CUR $w 16 ordered 3 >> gpurun_out/r03q_ab.jsonl 2>> gpurun_out/r03q_ab.err
  SPCU_AB_LANES=4 timeout 900 python profiles/scripts/ab_frame.py $V/libspcu_noilp.so,$CUR,$V/libspcu_noilp.so,$CUR $w 64 ordered 2 >> gpurun_out/r03q_ab.jsonl 2>> gpurun_out/r03q_ab.err
done
tail -n 3 gpurun_out/r03q_ab.err
