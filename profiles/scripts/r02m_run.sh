set -x
mkdir -p gpurun_out
V=build/variants
CUR=simplepath_b200/csrc/libspcu.so
for w in bunny_1080p_256spp:16 elf_1080p_256spp:16; do
  timeout 900 python profiles/scripts/ab_frame.py $V/libspcu_any_unsorted.so,$CUR,$V/libspcu_any_unsorted.so,$CUR ${w%%:*} ${w#*:} ordered 3 >> gpurun_out/r02m_ab_any.jsonl 2>> gpurun_out/r02m_ab.err
done
for i in 1 2 3; do
timeout 900 python profiles/scripts/ab_frame.py $V/libspcu_r1.so,$CUR,$V/libspcu_smw_oldrng.so,$V/libspcu_smw_nocaps.so,$V/libspcu_smw_both.so example_scene_1080p_64spp 64 default 5 >> gpurun_out/r02m_ab_smw.jsonl 2>> gpurun_out/r02m_ab.err
done
timeout 600 python -m pytest tests/test_gpu_trace.py -x -q -m gpu > gpurun_out/r02m_test_trace.log 2>&1
tail -n 2 gpurun_out/r02m_test_trace.log
