set -x
mkdir -p gpurun_out
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__thread_inst_executed.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active
for w in material_spheres_256_16spp:16:c1 example_scene_1080p_64spp:8:c2 elf_1080p_256spp:4:c4 lucy_4k_256spp:1:c5; do
  name=${w%%:*}; r=${w#*:}; spp=${r%%:*}; tag=${r#*:}
  timeout 1200 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02j_launches_$tag.csv python profiles/traffic_probe.py $name $spp auto ordered > gpurun_out/r02j_probe_$tag.json 2> gpurun_out/r02j_probe_$tag.err
done
timeout 1500 python profiles/dropin_build_probe.py c3_bunny c4_elf c5_lucy > gpurun_out/r02j_dropin_build.jsonl 2> gpurun_out/r02j_dropin_build.err
cat gpurun_out/r02j_dropin_build.jsonl; ls -la gpurun_out | grep r02j
