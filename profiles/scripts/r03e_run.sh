set -x
mkdir -p gpurun_out
N=$1
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r03e_bench_n$N.json 2> gpurun_out/r03e_bench_n$N.err ) 2> gpurun_out/r03e_bench_n$N.time
cat gpurun_out/r03e_bench_n$N.time; tail -n 3 gpurun_out/r03e_bench_n$N.err; cut -c1-250 gpurun_out/r03e_bench_n$N.json
timeout 900 python -m pytest tests/test_gpu_dropin.py -x -q -m gpu > gpurun_out/r03e_test_dropin_n$N.log 2>&1; tail -n 3 gpurun_out/r03e_test_dropin_n$N.log
