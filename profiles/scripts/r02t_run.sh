set -x
mkdir -p gpurun_out
V=build/variants
CUR=simplepath_b200/csrc/libspcu.so
timeout 1500 python -m pytest tests/test_gpu_trace.py tests/test_gpu_scale.py -x -q -m gpu > gpurun_out/r02t_test_trace.log 2>&1; echo "trace rc=$?"
tail -n 5 gpurun_out/r02t_test_trace.log
for w in bunny_1080p_256spp:16 elf_1080p_256spp:16; do
  timeout 900 python profiles/scripts/ab_frame.py $V/libspcu_noquads.so,$CUR,$V/libspcu_noquads.so,$CUR ${w%%:*} ${w#*:} ordered 3 >> gpurun_out/r02t_ab.jsonl 2>> gpurun_out/r02t_ab.err
done
tail -n 3 gpurun_out/r02t_ab.err
