set -x
mkdir -p gpurun_out
V=build/variants
CUR=simplepath_b200/csrc/libspcu.so
timeout 1500 python -m pytest tests/test_gpu_trace.py tests/test_gpu_scale.py tests/test_gpu_render.py -x -q -m gpu > gpurun_out/r03y_test.log 2>&1; echo "tests rc=$?"
tail -n 2 gpurun_out/r03y_test.log
for w in bunny_1080p_256spp elf_1080p_256spp; do
  SPCU_AB_LANES=1 timeout 900 python profiles/scripts/ab_frame.py $V/libspcu_serialrefill.so,$CUR,$V/libspcu_serialrefill.so,$CUR $w 16 ordered 3 >> gpurun_out/r03y_ab.jsonl 2>> gpurun_out/r03y_ab.err
  SPCU_AB_LANES=4 timeout 900 python profiles/scripts/ab_frame.py $V/libspcu_serialrefill.so,$CUR,$V/libspcu_serialrefill.so,$CUR $w 64 ordered 2 >> gpurun_out/r03y_ab.jsonl 2>> gpurun_out/r03y_ab.err
done
tail -n 3 gpurun_out/r03y_ab.err
