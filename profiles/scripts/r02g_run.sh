set -x
mkdir -p gpurun_out
for k in k_extend_walk:1 k_shadow_walk:1 k_extend_begin:1 k_shadow_begin:1; do
  name=${k%%:*}; skip=${k#*:}
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$name -s $skip -c 1 -f -o gpurun_out/r02g_$name python profiles/traffic_probe.py bunny_1080p_256spp 4 auto ordered > gpurun_out/r02g_ncu_$name.log 2>&1
done
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__thread_inst_executed.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active
timeout 900 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02g_launches_c3.csv python profiles/traffic_probe.py bunny_1080p_256spp 4 auto ordered > gpurun_out/r02g_probe_c3.json 2> gpurun_out/r02g_probe_c3.err
ls -la gpurun_out | grep r02g
