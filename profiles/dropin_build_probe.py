import os, subprocess, sys, time
sys.path.insert(0, '.')
from simplepath_b200 import host, scenes
for name, spp in (("c3_bunny", 4), ("c4_elf", 4), ("c5_lucy", 1)):
    sp = scenes.ensure(name)
    env = dict(os.environ, SPCU_BUILD_ON_DEVICE="1")
    t0 = time.time()
    p = subprocess.run([str(host.DRIVER), "--samples", str(spp), "--integrator", "cuda", sp.name], cwd=sp.parent, env=env,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    lines = [l for l in p.stdout.splitlines() if "CudaIntegrator" in l or "Elapsed" in l or "rror" in l]
    print(name, "rc", p.returncode, f"{time.time()-t0:.1f}s", lines[-4:], flush=True)
