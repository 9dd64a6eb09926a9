set -x
mkdir -p gpurun_out
V=build/variants
CUR=simplepath_b200/csrc/libspcu.so
timeout 900 python profiles/scripts/ab_frame.py $CUR,$V/libspcu_nopf.so,$V/libspcu_voteL.so,$V/libspcu_voteN.so,$V/libspcu_refill8.so,$V/libspcu_refill2.so,$CUR bunny_1080p_256spp 16 ordered 3 > gpurun_out/r02h_ab_c3.jsonl 2> gpurun_out/r02h_ab_c3.err
timeout 900 python profiles/scripts/ab_frame.py $CUR,$V/libspcu_nopf.so,$V/libspcu_voteL.so,$V/libspcu_voteN.so elf_1080p_256spp 16 ordered 3 > gpurun_out/r02h_ab_c4.jsonl 2> gpurun_out/r02h_ab_c4.err
( time timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/r02h_bench_default.json 2> gpurun_out/r02h_bench_default.err ) 2> gpurun_out/r02h_bench_default.time
tail -n 5 gpurun_out/r02h_bench_default.err; cat gpurun_out/r02h_bench_default.time
