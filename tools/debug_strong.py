"""Debug: exchange-build step vs solo step on one GPU (world = 1 endpoint: no peer, no wait), device loops."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import csv_simd_b200 as cs
from tools import gen

size = int(sys.argv[1]) if len(sys.argv) > 1 else (4 << 30)
dev = torch.device("cuda", 0)
ctx = cs.Context(0)
ex = ctx.exchange(0, 1)
data, _ = gen.quoted(size, seed=44)
n = data.size
d = torch.empty(n + 64, dtype=torch.uint8, device=dev)
d[:n].copy_(torch.from_numpy(data))
stream = torch.cuda.current_stream(dev)
ctx.set_stream(stream.cuda_stream)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for rnd in range(3):
    for name in ("exchange", "solo"):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(10):
            if name == "exchange":
                idx = ctx.index_build_shard_exchange(ex, d.data_ptr(), n, 0)
            else:
                idx = ctx.index_build_shard_device(d.data_ptr(), n, 0, 0, True)
            idx.free()
        e1.record(stream)
        th = time.perf_counter() - t0
        torch.cuda.synchronize()
        print(name, "ms/step", e0.elapsed_time(e1) / 10, "host enqueue ms/step", th * 100)
