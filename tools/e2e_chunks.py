"""Times the host-scalar C-ABI call for several chunkings (PLONKISH_CUDA_HOST_CHUNKS = equal chunks,
PLONKISH_CUDA_HOST_CUTS = chunk ends as fractions)."""
import os, sys, time, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    import numpy as np, torch
    import plonkish_b200 as pk
    log_n = int(sys.argv[2]); n = 1 << log_n
    host = torch.empty((n, 4), dtype=torch.int64).pin_memory()
    sc = host.numpy().view(np.uint64); sc[:] = pk.random_scalars(n, 1)
    d_bs = pk.synth_bases_device(n, 3, 5); torch.cuda.synchronize()
    reg = pk.G1Bases(d_bs)
    for _ in range(2): pk.variable_base_msm(sc, reg)
    t = time.perf_counter()
    for _ in range(5): out = pk.variable_base_msm(sc, reg)
    ms = (time.perf_counter() - t) / 5 * 1e3
    # plain H2D copy time for reference
    d = torch.empty_like(host, device="cuda"); torch.cuda.synchronize()
    t = time.perf_counter(); d.copy_(host); torch.cuda.synchronize(); cp = (time.perf_counter() - t) * 1e3
    print(json.dumps({"chunks": os.environ.get("PLONKISH_CUDA_HOST_CHUNKS"), "cuts": os.environ.get("PLONKISH_CUDA_HOST_CUTS"), "log_n": log_n, "ms": round(ms, 2), "h2d_ms": round(cp, 2)}))
else:
    for log_n in (24, 22, 20):
        subprocess.run([sys.executable, __file__, "child", str(log_n)], env=dict(os.environ))  # library default
        for cuts in ("1", "0.125,1", "0.25,1", "0.2,0.6,1", "0.125,0.5,1", "0.0625,0.25,0.625,1", "0.1,0.4,1", "0.125,0.4375,1", "0.0625,0.3125,1"):
            env = dict(os.environ, PLONKISH_CUDA_HOST_CUTS=cuts)
            subprocess.run([sys.executable, __file__, "child", str(log_n)], env=env)
