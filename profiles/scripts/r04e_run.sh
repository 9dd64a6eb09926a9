set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r04e_test_gpu.log 2>&1; echo "gpu tests rc=$?"
tail -n 2 gpurun_out/r04e_test_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r04e_smoke.log 2>&1; tail -n 1 gpurun_out/r04e_smoke.log
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__thread_inst_executed.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active
for w in bunny_1080p_256spp:4:c3 elf_1080p_256spp:4:c4 lucy_4k_256spp:1:c5; do
  name=${w%%:*}; r=${w#*:}; spp=${r%%:*}; tag=${r#*:}
  timeout 1200 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r04e_launches_$tag.csv python profiles/traffic_probe.py $name $spp auto ordered > gpurun_out/r04e_probe_$tag.json 2> gpurun_out/r04e_probe_$tag.err
done
( time timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r04e_bench_n1.json 2> gpurun_out/r04e_bench_n1.err ) 2> gpurun_out/r04e_bench_n1.time
cat gpurun_out/r04e_bench_n1.time; tail -n 2 gpurun_out/r04e_bench_n1.err; cut -c1-200 gpurun_out/r04e_bench_n1.json
