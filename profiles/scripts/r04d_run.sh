set -x
mkdir -p gpurun_out
V=build/variants
CUR=simplepath_b200/csrc/libspcu.so
timeout 1500 python -m pytest tests/test_gpu_trace.py tests/test_gpu_scale.py -x -q -m gpu > gpurun_out/r04d_test.log 2>&1; echo "tests rc=$?"
tail -n 2 gpurun_out/r04d_test.log
for w in bunny_1080p_256spp elf_1080p_256spp; do
  SPCU_AB_LANES=4 timeout 900 python profiles/scripts/ab_frame.py $CUR,$V/libspcu_v21.so,$V/libspcu_v12.so,$V/libspcu_v32.so,$V/libspcu_v23.so,$CUR $w 64 ordered 2 >> gpurun_out/r04d_ab.jsonl 2>> gpurun_out/r04d_ab.err
done
tail -n 3 gpurun_out/r04d_ab.err
