"""Call latency of `reader::read` at small and medium sizes: csvb200_index_build_to_host (host bytes in, host index out,
pinned and pageable) against the reference's SSE loop on one host core (oracle restatement), same bytes.  Where does the
GPU path start to win?  One JSON line per size; the index is compared with the oracle's at every size."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import csv_simd_b200 as cs  # noqa: E402
from oracle import oracle as O  # noqa: E402
from tools import gen  # noqa: E402


def best(fn, reps):
    t = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        t.append(time.perf_counter() - t0)
    return min(t), sorted(t)[len(t) // 2]


def main():
    ctx = cs.Context(0)
    full, _ = gen.quoted(256 << 20, seed=43)
    for size in (4 << 10, 64 << 10, 1 << 20, 4 << 20, 16 << 20, 64 << 20, 256 << 20):
        data = full[:size] if size < full.size else full
        n = int(data.size)
        want = O.read_sse(data)
        h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
        h_in.numpy()[:] = data
        h_out = torch.zeros(want.size + 64, dtype=torch.int64).pin_memory()
        p_in = np.array(data, copy=True)
        p_out = np.zeros(want.size + 64, dtype=np.uint64)
        reps = 30 if size <= (16 << 20) else 8
        for _ in range(3):
            ln = ctx.index_build_to_host(h_in.data_ptr(), n, h_out.data_ptr(), h_out.numel())
            ctx.index_build_to_host(p_in.ctypes.data, n, p_out.ctypes.data, p_out.size)
        ok = bool(ln == want.size and (h_out.numpy()[:ln].view(np.uint64) == want).all() and (p_out[:ln] == want).all())
        pin_min, pin_med = best(lambda: ctx.index_build_to_host(h_in.data_ptr(), n, h_out.data_ptr(), h_out.numel()), reps)
        pg_min, pg_med = best(lambda: ctx.index_build_to_host(p_in.ctypes.data, n, p_out.ctypes.data, p_out.size), reps)
        al = O.aligned_copy(data)
        cpu_min, cpu_med = best(lambda: O.read_sse_timed(al), max(3, min(reps, (64 << 20) // max(n, 1))))
        print(json.dumps({"bytes": n, "index_entries": int(want.size), "parity_ok": ok,
                          "gpu_pinned_us": round(pin_med * 1e6, 1), "gpu_pageable_us": round(pg_med * 1e6, 1),
                          "cpu_reference_1core_us": round(cpu_med * 1e6, 1),
                          "gpu_pinned_gbs": round(n / pin_med / 1e9, 2), "gpu_pageable_gbs": round(n / pg_med / 1e9, 2),
                          "cpu_gbs": round(n / cpu_med / 1e9, 2), "speedup_pinned": round(cpu_med / pin_med, 2),
                          "speedup_pageable": round(cpu_med / pg_med, 2)}), flush=True)
        assert ok
    ctx.close()


if __name__ == "__main__":
    main()
