set -x
mkdir -p gpurun_out
V=build/variants
CUR=simplepath_b200/csrc/libspcu.so
timeout 1200 python -m pytest tests/test_gpu_trace.py tests/test_gpu_scale.py -x -q -m gpu > gpurun_out/r02l_test_trace.log 2>&1; echo "trace rc=$?"
timeout 1500 python -m pytest tests/test_gpu_render.py -x -q -m gpu > gpurun_out/r02l_test_render.log 2>&1; echo "render rc=$?"
for w in bunny_1080p_256spp:16 elf_1080p_256spp:16; do
  timeout 900 python profiles/scripts/ab_frame.py $V/libspcu_mis_single.so,$CUR,$V/libspcu_mis_single.so,$CUR ${w%%:*} ${w#*:} ordered 3 >> gpurun_out/r02l_ab.jsonl 2>> gpurun_out/r02l_ab.err
done
timeout 900 python profiles/scripts/ab_frame.py $V/libspcu_r1.so,$CUR,$V/libspcu_r1.so,$CUR example_scene_1080p_64spp 64 default 5 >> gpurun_out/r02l_ab.jsonl 2>> gpurun_out/r02l_ab.err
tail -n 3 gpurun_out/r02l_test_*.log; cut -c1-420 gpurun_out/r02l_ab.jsonl
