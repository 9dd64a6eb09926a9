set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02i_bench_n2.json 2> gpurun_out/r02i_bench_n2.err; echo "bench n2 rc=$?"
timeout 900 python -m pytest tests/test_gpu_dropin.py -x -q -m gpu > gpurun_out/r02i_test_dropin.log 2>&1; echo "dropin rc=$?"
tail -n 15 gpurun_out/r02i_test_dropin.log; tail -n 5 gpurun_out/r02i_bench_n2.err; cut -c1-600 gpurun_out/r02i_bench_n2.json
