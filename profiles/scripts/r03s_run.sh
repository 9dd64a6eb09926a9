set -x
mkdir -p gpurun_out
N=$1
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r03s_bench_n$N.json 2> gpurun_out/r03s_bench_n$N.err ) 2> gpurun_out/r03s_bench_n$N.time
cat gpurun_out/r03s_bench_n$N.time; tail -n 3 gpurun_out/r03s_bench_n$N.err; cut -c1-250 gpurun_out/r03s_bench_n$N.json
