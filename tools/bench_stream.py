"""SURVEY 8f rank 3 measured: file -> host index through csvb200_index_build_file (thread-pool pread into
the pinned ring, chained launches, overlapped D2H), next to the reference's path on the same file (mmap +
the single-threaded SSE loop, page faults included -- that is what csv_simd::create does, src/lib.rs:61-74).

    python tools/bench_stream.py [--size BYTES] [--dir DIR]
"""
import argparse
import json
import mmap
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=4 << 30)
    ap.add_argument("--dir", default="/dev/shm" if os.path.isdir("/dev/shm") else "/tmp")
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    from csv_simd_b200 import numa
    numa.bind_to_device(0)
    path = os.path.join(a.dir, "csvb200_stream_bench.csv")
    data, rows = gen.unquoted(a.size, seed=42)
    n = data.size
    data.tofile(path)
    ctx = cs.Context(0)
    try:
        E_guess = n // 3 + 4096
        h = torch.empty(E_guess, dtype=torch.int64).pin_memory()
        res = {}
        ln = 0
        for label, run in (("pinned_dst", lambda: ctx.index_build_file_ptr(path, h.data_ptr(), h.numel())),
                           ("pageable_dst", None)):
            if run is None:
                out = np.empty(E_guess, dtype=np.uint64)
                out[::512] = 0    # fault the pages in once, as a reused Vec would be
                run = lambda: ctx.index_build_file_ptr(path, out.ctypes.data, out.size)  # noqa: E731
            run()
            best = 1e9
            for _ in range(a.reps):
                t = time.perf_counter()
                ln, st = run()
                best = min(best, time.perf_counter() - t)
            res[label] = {"value": n / best / 1e9, "unit": "GB/s", "seconds": best, "chunks": st["chunks"]}
        # parity: checksum of the whole index against the oracle run over the mmap'd file (also the CPU baseline)
        got = h.numpy()[:ln].view(np.uint64)
        with open(path, "rb") as f:
            mm = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ)
            buf = np.frombuffer(mm, dtype=np.uint8)
            t = time.perf_counter()
            E_cpu, cs_cpu = O.read_sse_timed(buf)
            cpu_s = time.perf_counter() - t
            del buf
            mm.close()
        assert E_cpu == ln, (E_cpu, ln)
        assert int(got.sum(dtype=np.uint64)) == cs_cpu, "index checksum differs from the oracle's"
        assert bool((np.diff(got[1:].view(np.int64)) > 0).all()), "index not strictly increasing"
        line = {
            "metric": "csv_file_bytes_indexed_per_sec", "unit": "GB/s", "value": res["pinned_dst"]["value"],
            "config": {"workload": "cfg2_unquoted file", "file_bytes": int(n), "dir": a.dir, "index_entries": int(ln),
                       "host_cpus": os.cpu_count()},
            "file_to_pinned_index": res["pinned_dst"], "file_to_pageable_index": res["pageable_dst"],
            "h2d_bytes": int(n), "d2h_bytes": int(8 * ln),
            "cpu_baseline": {"value": n / cpu_s / 1e9, "unit": "GB/s", "cores": 1, "kind": "port",
                             "sample": "whole file through mmap + the SSE restatement, single thread, page faults "
                                       "included (what csv_simd::create does)"},
            "parity": "entry count and wrapping sum of all entries equal the oracle's; strictly increasing",
        }
        print(json.dumps(line), flush=True)
    finally:
        ctx.close()
        os.unlink(path)


if __name__ == "__main__":
    main()
