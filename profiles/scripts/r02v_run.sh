set -x
mkdir -p gpurun_out
V=build/variants
CUR=simplepath_b200/csrc/libspcu.so
timeout 1500 python -m pytest tests/test_gpu_render.py tests/test_gpu_converged.py -x -q -m gpu > gpurun_out/r02v_test_render.log 2>&1; echo "render rc=$?"
tail -n 5 gpurun_out/r02v_test_render.log
for w in bunny_1080p_256spp:16 elf_1080p_256spp:16; do
  timeout 900 python profiles/scripts/ab_frame.py $V/libspcu_nofuse.so,$CUR,$V/libspcu_nofuse.so,$CUR ${w%%:*} ${w#*:} ordered 3 >> gpurun_out/r02v_ab.jsonl 2>> gpurun_out/r02v_ab.err
done
tail -n 3 gpurun_out/r02v_ab.err
