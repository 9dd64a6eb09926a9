set -x
mkdir -p gpurun_out
V=build/variants
CUR=simplepath_b200/csrc/libspcu.so
timeout 1500 python -m pytest tests/test_gpu_trace.py tests/test_gpu_scale.py -x -q -m gpu > gpurun_out/r02u_test_trace.log 2>&1; echo "trace rc=$?"
tail -n 5 gpurun_out/r02u_test_trace.log
for w in bunny_1080p_256spp:16 elf_1080p_256spp:16; do
  timeout 900 python profiles/scripts/ab_frame.py $V/libspcu_prim128.so,$CUR,$V/libspcu_stk8.so,$V/libspcu_stk12.so,$V/libspcu_prim128.so,$CUR ${w%%:*} ${w#*:} ordered 3 >> gpurun_out/r02u_ab.jsonl 2>> gpurun_out/r02u_ab.err
done
tail -n 3 gpurun_out/r02u_ab.err
