mkdir -p gpurun_out
SPCU_DEBUG_SYNC=1 timeout 300 python profiles/scripts/sanitize_pipelines.py g_spheres wavefront iterative_rrnee ordered shadow > gpurun_out/r02e_pipelines3.log 2>&1
grep -n "SPCU_DEBUG\|ok" gpurun_out/r02e_pipelines3.log
