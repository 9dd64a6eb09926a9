set -x
mkdir -p gpurun_out
( while true; do nvidia-smi --query-gpu=memory.used,clocks.sm,temperature.gpu --format=csv,noheader; sleep 10; done ) > gpurun_out/r03r_smi.log 2>&1 &
SMI=$!
( time timeout 1500 python bench.py --steps 100 --warmup 5 --no-cpu --no-side-configs > gpurun_out/r03r_bench_n1_steps100.json 2> gpurun_out/r03r_bench.err ) 2> gpurun_out/r03r_bench.time
kill $SMI
tail -n 2 gpurun_out/r03r_bench.err; cat gpurun_out/r03r_bench.time; cat gpurun_out/r03r_smi.log | tr '\n' ';' | cut -c1-600
