set -x
mkdir -p gpurun_out
V=build/variants
CUR=simplepath_b200/csrc/libspcu.so
for w in bunny_1080p_256spp:16 elf_1080p_256spp:16; do
  timeout 900 python profiles/scripts/ab_frame.py $CUR,$V/libspcu_begin12.so,$V/libspcu_begin16.so,$CUR ${w%%:*} ${w#*:} ordered 3 >> gpurun_out/r02w_ab.jsonl 2>> gpurun_out/r02w_ab.err
done
tail -n 3 gpurun_out/r02w_ab.err
