set -x
mkdir -p gpurun_out
V=build/variants
CUR=simplepath_b200/csrc/libspcu.so
timeout 1200 python -m pytest tests/test_gpu_trace.py tests/test_gpu_scale.py -x -q -m gpu > gpurun_out/r02c_test_trace.log 2>&1; echo "trace rc=$?"
timeout 1500 python -m pytest tests/test_gpu_render.py -x -q -m gpu > gpurun_out/r02c_test_render.log 2>&1; echo "render rc=$?"
timeout 900 python profiles/scripts/ab_frame.py $V/libspcu_r1.so,$CUR,$V/libspcu_mb8.so,$V/libspcu_nopair_mb8.so,$V/libspcu_ld128_mb8.so,$V/libspcu_rho_literal.so bunny_1080p_256spp 16 ordered 3 > gpurun_out/r02c_ab_c3_ordered.jsonl 2> gpurun_out/r02c_ab_c3_ordered.err
timeout 900 python profiles/scripts/ab_frame.py $V/libspcu_r1.so,$CUR,$V/libspcu_mb8.so,$V/libspcu_nopair_mb8.so,$V/libspcu_ld128_mb8.so elf_1080p_256spp 16 ordered 3 > gpurun_out/r02c_ab_c4_ordered.jsonl 2> gpurun_out/r02c_ab_c4_ordered.err
timeout 900 python profiles/scripts/ab_frame.py $V/libspcu_r1.so,$CUR,$V/libspcu_smw_oldrng.so,$V/libspcu_smw_nocaps.so,$V/libspcu_r1.so,$CUR example_scene_1080p_64spp 64 default 5 > gpurun_out/r02c_ab_c2.jsonl 2> gpurun_out/r02c_ab_c2.err
timeout 900 python -m pytest tests/test_gpu_converged.py -x -q -m gpu -s > gpurun_out/r02c_test_converged.log 2>&1; echo "converged rc=$?"
tail -n 3 gpurun_out/r02c_test_*.log
