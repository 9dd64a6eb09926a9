set -x
mkdir -p gpurun_out
CUR=simplepath_b200/csrc/libspcu.so
timeout 1500 python -m pytest tests/test_gpu_render.py tests/test_gpu_converged.py tests/test_gpu_dropin.py -x -q -m gpu > gpurun_out/r02z_test_render.log 2>&1; echo "render rc=$?"
tail -n 5 gpurun_out/r02z_test_render.log
for w in bunny_1080p_256spp elf_1080p_256spp; do
  for lanes in 1 2 3 4 6; do
    SPCU_AB_LANES=$lanes timeout 900 python profiles/scripts/ab_frame.py $CUR $w 64 ordered 2 >> gpurun_out/r02z_ab.jsonl 2>> gpurun_out/r02z_ab.err
  done
done
SPCU_AB_LANES=1 timeout 900 python profiles/scripts/ab_frame.py $CUR bunny_1080p_256spp 16 ordered 2 >> gpurun_out/r02z_ab.jsonl 2>> gpurun_out/r02z_ab.err
SPCU_AB_LANES=2 timeout 900 python profiles/scripts/ab_frame.py $CUR bunny_1080p_256spp 16 ordered 2 >> gpurun_out/r02z_ab.jsonl 2>> gpurun_out/r02z_ab.err
tail -n 3 gpurun_out/r02z_ab.err
