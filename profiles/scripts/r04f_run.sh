set -x
mkdir -p gpurun_out
N=$1
if [ "$N" = "1" ]; then
  ( time timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r04f_bench_n1.json 2> gpurun_out/r04f_bench_n1.err ) 2> gpurun_out/r04f_bench_n1.time
else
  ( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r04f_bench_n$N.json 2> gpurun_out/r04f_bench_n$N.err ) 2> gpurun_out/r04f_bench_n$N.time
fi
cat gpurun_out/r04f_bench_n$N.time; tail -n 2 gpurun_out/r04f_bench_n$N.err; cut -c1-200 gpurun_out/r04f_bench_n$N.json
