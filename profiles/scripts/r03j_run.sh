set -x
mkdir -p gpurun_out
CUR=simplepath_b200/csrc/libspcu.so
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r03j_test_gpu.log 2>&1; echo "gpu tests rc=$?"
tail -n 3 gpurun_out/r03j_test_gpu.log
for lanes in 1 2 4 8; do
  SPCU_AB_LANES=$lanes timeout 900 python profiles/scripts/ab_frame.py $CUR material_spheres_256_16spp 16 ordered 20 >> gpurun_out/r03j_ab.jsonl 2>> gpurun_out/r03j_ab.err
done
tail -n 3 gpurun_out/r03j_ab.err
