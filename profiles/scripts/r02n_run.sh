mkdir -p gpurun_out
V=build/variants
CUR=simplepath_b200/csrc/libspcu.so
for i in 1 2; do
timeout 900 python profiles/scripts/ab_frame.py $V/libspcu_r1.so,$CUR,$V/libspcu_smw_rngcall.so example_scene_1080p_64spp 64 default 5 >> gpurun_out/r02n_ab_smw.jsonl 2>> gpurun_out/r02n_ab.err
done
cut -c1-330 gpurun_out/r02n_ab_smw.jsonl
