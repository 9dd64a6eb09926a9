set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02a_gpus.txt 2>&1
timeout 1200 python -m pytest tests/test_gpu_trace.py -x -q -m gpu > gpurun_out/r02a_test_trace.log 2>&1; echo "trace rc=$?"
timeout 900 python -m pytest tests/test_gpu_scale.py -x -q -m gpu -s > gpurun_out/r02a_test_scale.log 2>&1; echo "scale rc=$?"
timeout 1500 python -m pytest tests/test_gpu_render.py -x -q -m gpu > gpurun_out/r02a_test_render.log 2>&1; echo "render rc=$?"
for t in exact ordered; do
  timeout 600 python bench.py --workload bunny_1080p_256spp --spp 16 --steps 3 --warmup 3 --no-cpu --traversal $t > gpurun_out/r02a_bench_c3_$t.json 2> gpurun_out/r02a_bench_c3_$t.err; echo "bench $t rc=$?"
done
timeout 600 python bench.py --workload elf_1080p_256spp --spp 16 --steps 3 --warmup 3 --no-cpu --traversal ordered > gpurun_out/r02a_bench_c4_ordered.json 2> gpurun_out/r02a_bench_c4_ordered.err; echo "bench c4 rc=$?"
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r02a_bench_c2.json 2> gpurun_out/r02a_bench_c2.err; echo "bench c2 rc=$?"
tail -3 gpurun_out/r02a_test_*.log
