set -x
mkdir -p gpurun_out
V=build/variants
CUR=simplepath_b200/csrc/libspcu.so
timeout 900 python profiles/scripts/ab_frame.py $V/libspcu_r1.so,$CUR,$V/libspcu_mb8.so,$V/libspcu_mb10.so,$V/libspcu_mb12.so,$V/libspcu_nopair.so,$V/libspcu_nopair_mb8.so,$V/libspcu_noslab_mb8.so bunny_1080p_256spp 16 ordered 3 > gpurun_out/r02b_ab_c3_ordered.jsonl 2> gpurun_out/r02b_ab_c3_ordered.err
timeout 900 python profiles/scripts/ab_frame.py $V/libspcu_r1.so,$CUR,$V/libspcu_mb8.so,$V/libspcu_nopair_mb8.so bunny_1080p_256spp 16 exact 3 > gpurun_out/r02b_ab_c3_exact.jsonl 2> gpurun_out/r02b_ab_c3_exact.err
timeout 900 python profiles/scripts/ab_frame.py $CUR,$V/libspcu_nee6.so,$V/libspcu_nee5.so,$V/libspcu_nee4.so,$V/libspcu_shade6.so bunny_1080p_256spp 16 ordered 3 > gpurun_out/r02b_ab_c3_shade.jsonl 2> gpurun_out/r02b_ab_c3_shade.err
timeout 900 python profiles/scripts/ab_frame.py $V/libspcu_r1.so,$CUR,$V/libspcu_r1.so,$CUR example_scene_1080p_64spp 64 default 5 > gpurun_out/r02b_ab_c2.jsonl 2> gpurun_out/r02b_ab_c2.err
timeout 900 python profiles/scripts/ab_frame.py $V/libspcu_r1.so,$CUR,$V/libspcu_mb8.so,$V/libspcu_nopair_mb8.so elf_1080p_256spp 16 ordered 3 > gpurun_out/r02b_ab_c4_ordered.jsonl 2> gpurun_out/r02b_ab_c4_ordered.err
for k in k_extend_walk:1 k_shadow_walk:1 k_nee_bsdf:0 k_shade:1; do
  name=${k%%:*}; skip=${k#*:}
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$name -s $skip -c 1 -f -o gpurun_out/r02b_$name python profiles/traffic_probe.py bunny_1080p_256spp 4 auto ordered > gpurun_out/r02b_ncu_$name.log 2>&1
done
ls -la gpurun_out
