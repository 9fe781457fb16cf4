"""How long does cudaHostRegister / cudaHostUnregister of a pageable buffer take (the alternative to the staging ring)?"""
import time, sys, os
import numpy as np, torch
rt = torch.cuda.cudart()
torch.cuda.init()
d = torch.empty(1 << 29, dtype=torch.uint8, device="cuda")
print("nproc", os.cpu_count())
for mb in (16, 64, 512):
    a = np.random.default_rng(1).integers(0, 255, size=mb << 20, dtype=np.uint8)
    ptr = a.ctypes.data
    for rep in range(3):
        t0 = time.perf_counter(); r = rt.cudaHostRegister(ptr, a.nbytes, 0); t1 = time.perf_counter()
        t = torch.from_numpy(a)
        torch.cuda.synchronize(); t2 = time.perf_counter()
        d[: a.nbytes].copy_(t, non_blocking=True); torch.cuda.synchronize(); t3 = time.perf_counter()
        u = rt.cudaHostUnregister(ptr); t4 = time.perf_counter()
        print(f"{mb} MB rep {rep}: register {1e3*(t1-t0):.2f} ms ({r}), copy {1e3*(t3-t2):.2f} ms, unregister {1e3*(t4-t3):.2f} ms ({u})", flush=True)
