set -x
mkdir -p gpurun_out
( time timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02p_bench_n1.json 2> gpurun_out/r02p_bench_n1.err ) 2> gpurun_out/r02p_bench_n1.time
( time timeout 1500 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02p_bench_ref_n1.json 2> gpurun_out/r02p_bench_ref_n1.err ) 2> gpurun_out/r02p_bench_ref_n1.time
cat gpurun_out/r02p_bench_n1.time gpurun_out/r02p_bench_ref_n1.time; cut -c1-300 gpurun_out/r02p_bench_ref_n1.json
