set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r03p_test_gpu.log 2>&1; echo "gpu tests rc=$?"
tail -n 3 gpurun_out/r03p_test_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r03p_smoke.log 2>&1; tail -n 2 gpurun_out/r03p_smoke.log
( time timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r03p_bench_n1.json 2> gpurun_out/r03p_bench_n1.err ) 2> gpurun_out/r03p_bench_n1.time
cat gpurun_out/r03p_bench_n1.time; tail -n 3 gpurun_out/r03p_bench_n1.err
