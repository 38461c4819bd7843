"""Where the pageable end-to-end time goes: csvb200_index_build_to_host on the 1 GiB cfg2 workload with each side
(input bytes, output index) either pinned or ordinary pageable memory.  One JSON line per combination."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import csv_simd_b200 as cs  # noqa: E402
from tools import gen  # noqa: E402


def main():
    size = int(sys.argv[1]) if len(sys.argv) > 1 else (1 << 30)
    ctx = cs.Context(0)
    data, _ = gen.unquoted(size, seed=42)
    n = int(data.size)
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_in.numpy()[:] = data
    cap = n // 4 + 1024
    h_out = torch.empty(cap, dtype=torch.int64).pin_memory()
    p_out = np.empty(cap, dtype=np.uint64)
    p_out[::512] = 0
    for name, src, dst in (("pinned/pinned", h_in.data_ptr(), h_out.data_ptr()), ("pageable/pinned", data.ctypes.data, h_out.data_ptr()),
                           ("pinned/pageable", h_in.data_ptr(), p_out.ctypes.data), ("pageable/pageable", data.ctypes.data, p_out.ctypes.data)):
        ctx.index_build_to_host(src, n, dst, cap)
        ts = []
        for _ in range(4):
            t0 = time.perf_counter()
            ln = ctx.index_build_to_host(src, n, dst, cap)
            ts.append(time.perf_counter() - t0)
        print(json.dumps({"in/out": name, "bytes": n, "entries": int(ln), "ms_best": min(ts) * 1e3, "ms_all": [round(t * 1e3, 2) for t in ts],
                          "gbs": n / min(ts) / 1e9, "io_threads": os.environ.get("CSVB200_IO_THREADS")}), flush=True)
    # the host copy alone: numpy (one thread) and the library's slices are not exported, so time a plain copy
    t0 = time.perf_counter()
    p_out2 = data.copy()
    print(json.dumps({"one_thread_copy_gbs": n / (time.perf_counter() - t0) / 1e9, "note": "np.copy into fresh pages (page faults included)"}))
    t0 = time.perf_counter()
    p_out2[:] = data
    print(json.dumps({"one_thread_copy_gbs": n / (time.perf_counter() - t0) / 1e9, "note": "np copy into touched pages"}))
    ctx.close()


if __name__ == "__main__":
    main()
