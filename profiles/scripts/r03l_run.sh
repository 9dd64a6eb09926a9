set -x
mkdir -p gpurun_out
SPCU_LIB=$PWD/build/variants/libspcu_ieee.so timeout 1500 python -m pytest tests/test_gpu_render.py tests/test_gpu_converged.py -x -q -m gpu > gpurun_out/r03l_test_render_ieee_build.log 2>&1; echo "ieee build render tests rc=$?"
tail -n 4 gpurun_out/r03l_test_render_ieee_build.log
SPCU_AB_LANES=4 timeout 900 python profiles/scripts/ab_frame.py build/variants/libspcu_ieee.so,simplepath_b200/csrc/libspcu.so bunny_1080p_256spp 32 ordered 2 >> gpurun_out/r03l_ab_ieee.jsonl 2>> gpurun_out/r03l_ab.err
