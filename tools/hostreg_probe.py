"""cudaHostRegister / cudaHostUnregister throughput on ordinary (touched) pageable memory, whole and in 64 MiB pieces."""
import json, time, sys
import numpy as np, torch
rt = torch.cuda.cudart()
torch.cuda.init()
for mb in (64, 1024):
    a = np.empty(mb << 20, dtype=np.uint8)
    a[::4096] = 1
    for rep in range(2):
        t0 = time.perf_counter(); r = rt.cudaHostRegister(a.ctypes.data, a.nbytes, 0); t1 = time.perf_counter()
        u = rt.cudaHostUnregister(a.ctypes.data); t2 = time.perf_counter()
        print(json.dumps({"mb": mb, "rep": rep, "register_ms": (t1 - t0) * 1e3, "unregister_ms": (t2 - t1) * 1e3,
                          "register_gbs": a.nbytes / (t1 - t0) / 1e9, "rc": [int(r), int(u)]}))
