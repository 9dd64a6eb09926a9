set -x
mkdir -p gpurun_out
V=build/variants
CUR=simplepath_b200/csrc/libspcu.so
timeout 1200 python -m pytest tests/test_gpu_trace.py tests/test_gpu_scale.py -x -q -m gpu -s > gpurun_out/r02d_test_trace.log 2>&1; echo "trace rc=$?"
timeout 1500 python -m pytest tests/test_gpu_render.py -x -q -m gpu > gpurun_out/r02d_test_render.log 2>&1; echo "render rc=$?"
timeout 900 python profiles/scripts/ab_frame.py $V/libspcu_mb8.so,$CUR,$V/libspcu_mb8.so,$CUR bunny_1080p_256spp 16 ordered 3 > gpurun_out/r02d_ab_c3_ordered.jsonl 2> gpurun_out/r02d_ab_c3_ordered.err
timeout 900 python profiles/scripts/ab_frame.py $V/libspcu_mb8.so,$CUR elf_1080p_256spp 16 ordered 3 > gpurun_out/r02d_ab_c4_ordered.jsonl 2> gpurun_out/r02d_ab_c4_ordered.err
tail -n 3 gpurun_out/r02d_test_*.log
