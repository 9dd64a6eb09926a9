set -x
mkdir -p gpurun_out
S=profiles/scripts/sanitize_pipelines.py
L=gpurun_out/r03m_sanitizer.log
: > $L
for tool in memcheck racecheck; do
  for scene in g_elf g_bunny g_lights; do
    echo "== $tool $scene wavefront iterative_rrnee ordered, 600-slot batches (4 lanes)" >> $L
    SPCU_SAN_WAVEFRONT=600 timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 python $S $scene wavefront iterative_rrnee ordered >> $L 2>&1; echo "rc=$?" >> $L
  done
done
echo "== memcheck g_elf extend stage on ray batches (ordered)" >> $L
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python $S g_elf wavefront iterative_rrnee ordered extend >> $L 2>&1; echo "rc=$?" >> $L
echo "== memcheck g_elf shadow stage on ray batches" >> $L
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python $S g_elf wavefront iterative_rrnee ordered shadow >> $L 2>&1; echo "rc=$?" >> $L
grep -n "^==\|rc=\|ERROR SUMMARY\|RACECHECK SUMMARY\| ok " $L
