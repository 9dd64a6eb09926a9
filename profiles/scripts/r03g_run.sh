set -x
mkdir -p gpurun_out
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__thread_inst_executed.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active
for w in bunny_1080p_256spp:4:c3 material_spheres_256_16spp:16:c1 example_scene_1080p_64spp:8:c2 elf_1080p_256spp:4:c4 lucy_4k_256spp:1:c5; do
  name=${w%%:*}; r=${w#*:}; spp=${r%%:*}; tag=${r#*:}
  timeout 1200 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r03g_launches_$tag.csv python profiles/traffic_probe.py $name $spp auto ordered > gpurun_out/r03g_probe_$tag.json 2> gpurun_out/r03g_probe_$tag.err
done
for k in k_extend_walk:1 k_shadow_walk:1 k_extend_begin:1 k_nee_bsdf:1 k_shade:1; do
  name=${k%%:*}; skip=${k#*:}
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$name -s $skip -c 1 -f -o gpurun_out/r03g_$name python profiles/traffic_probe.py bunny_1080p_256spp 4 auto ordered > gpurun_out/r03g_ncu_$name.log 2>&1
done
ls -la gpurun_out | grep r03g
