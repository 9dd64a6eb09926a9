set -x
mkdir -p gpurun_out
( time timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r03h_bench_n1.json 2> gpurun_out/r03h_bench_n1.err ) 2> gpurun_out/r03h_bench_n1.time
( time timeout 1500 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r03h_bench_ref_n1.json 2> gpurun_out/r03h_bench_ref_n1.err ) 2> gpurun_out/r03h_bench_ref_n1.time
cat gpurun_out/r03h_bench_n1.time gpurun_out/r03h_bench_ref_n1.time; cut -c1-300 gpurun_out/r03h_bench_ref_n1.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r03h_launches_default_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-side-configs > gpurun_out/r03h_ncu_bench.log 2>&1
