"""All GPUs of the box from ONE process through csvb200_multi_* (no torch.distributed, no NCCL):

  1. config 4 end to end: one host buffer of G GiB quote-heavy CSV -> ONE contiguous host index
     (csvb200_multi_index_build_to_host: one host thread per GPU, exchange over peer-mapped mailboxes);
  2. config 5 at G GPUs: the index of a 4 GiB 256-field file stays distributed (segment k on GPU k), 10 M random
     (record, field) lookups are split over the GPUs and every kernel reads remote slots over NVLink
     (csvb200_multi_seek_fields / _device).

Parity: (1) entry count + wrapping sum + first / last 4096 entries against the oracle, (2) hit count and checksum of every
(start, end) pair against the oracle's seek_field.  One JSON line per measurement.

    python tools/bench_multi.py [--gpus G] [--gib-per-gpu 1] [--lookup-gib 4]
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import csv_simd_b200 as cs  # noqa: E402
from oracle import oracle as O  # noqa: E402
from tools import gen  # noqa: E402

GiB = 1 << 30


def gen_parallel(pieces, fn):
    out = [None] * pieces
    ts = [threading.Thread(target=lambda j=j: out.__setitem__(j, fn(j))) for j in range(pieces)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=torch.cuda.device_count())
    ap.add_argument("--gib-per-gpu", type=float, default=1.0)
    ap.add_argument("--lookup-gib", type=float, default=4.0)
    ap.add_argument("--queries", type=int, default=10_000_000)
    ap.add_argument("--steps", type=int, default=5)
    a = ap.parse_args()
    G = a.gpus
    devs = list(range(G)) if torch.cuda.device_count() >= G else [0] * G
    m = cs.Multi(devs)
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0

    # ---- 1. config 4 end to end through one call --------------------------------------------------------------
    per = int(a.gib_per_gpu * GiB)
    parts = gen_parallel(G, lambda j: gen.quoted(per, seed=44, first_row=j << 32, with_header=(j == 0))[0])
    n = int(sum(p.size for p in parts))
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    off = 0
    for p in parts:
        h_in[off:off + p.size].copy_(torch.from_numpy(p))
        off += p.size
    del parts
    data = h_in.numpy()
    t = time.perf_counter()
    cnt, checksum = O.read_sse_timed(O.aligned_copy(data)) if n <= 2 * GiB else (None, None)
    if cnt is None:   # the oracle over the whole file, piecewise (one thread; kept out of the timed region)
        want, _ = O.read_closed_form(data, 0, 0, with_sentinel=True)
        cnt, checksum = int(want.size), int(want.sum(dtype=np.uint64))
        head, tail = want[:4096].copy(), want[-4096:].copy()
        del want
    else:
        want = O.read_sse(data)
        head, tail = want[:4096].copy(), want[-4096:].copy()
        del want
    oracle_s = time.perf_counter() - t
    h_out = torch.empty(cnt + 1024, dtype=torch.int64).pin_memory()
    for _ in range(2):
        ln = m.index_build_to_host(h_in.data_ptr(), n, h_out.data_ptr(), h_out.numel())
    best, stats = 1e9, None
    for _ in range(a.steps):
        t = time.perf_counter()
        ln = m.index_build_to_host(h_in.data_ptr(), n, h_out.data_ptr(), h_out.numel())
        dt = time.perf_counter() - t
        if dt < best:
            best, stats = dt, m.stats()
    got = h_out.numpy()[:ln].view(np.uint64)
    ok = bool(ln == cnt and int(got.sum(dtype=np.uint64)) == checksum and (got[:4096] == head).all() and (got[-4096:] == tail).all())
    print(json.dumps({"bench": "multi_index_build_to_host", "api": "csvb200_multi_index_build_to_host (one process, one host thread per GPU)",
                      "gpus": G, "devices": devs, "workload": "cfg4 grammar, one host buffer cut at k*n/G + 37k + 13",
                      "bytes": n, "index_entries": int(ln), "seconds_best": best, "value": n / best / 1e9, "unit": "GB/s",
                      "upload_phase_s": stats["upload_seconds"], "download_phase_s": stats["download_seconds"],
                      "carry_mask": stats["carry_mask"], "redone_mask": stats["redone_mask"],
                      "h2d_bytes": n, "d2h_bytes": 8 * int(ln),
                      "parity": {"checked": True, "ok": ok, "what": "entry count, wrapping sum, first and last 4096 entries == oracle"},
                      "oracle_seconds_one_thread": oracle_s}), flush=True)
    assert ok
    del h_in, h_out, data

    # ---- 2. config 5 over a distributed index ----------------------------------------------------------------------
    size = int(a.lookup_gib * GiB)
    data, rows = gen.unquoted(size, seed=45, nfields=256, modulus=10 ** 15)
    n = int(data.size)
    t = time.perf_counter()
    mi = m.index_build_distributed(data)
    build_s = time.perf_counter() - t
    rc, jump = mi.tape_init(256, False)
    assert rc == rows + 1 and jump == 256
    nq = a.queries
    rec, fld = gen.queries(nq, rc, 256, seed=46)
    host = mi.to_host()
    cs_cpu, hits = O.seek_fields_timed(host, n, rc, 256, False, rec, fld)
    # host arrays in / out (pinned): every GPU takes nq / G queries
    h_rec, h_fld = torch.from_numpy(rec.view(np.int32)).pin_memory(), torch.from_numpy(fld.view(np.int32)).pin_memory()
    h_res = torch.empty((nq, 2), dtype=torch.int64).pin_memory()
    import ctypes as C
    best = 1e9
    for _ in range(3):
        t = time.perf_counter()
        rcode = mi._lib.csvb200_multi_seek_fields(mi._h, C.c_void_p(h_rec.data_ptr()), C.c_void_p(h_fld.data_ptr()), nq,
                                                  C.c_void_p(h_res.data_ptr()))
        best = min(best, time.perf_counter() - t)
        assert rcode == 0
    got = h_res.numpy().view(np.uint64)
    live = got[:, 0] != np.uint64(0xFFFFFFFFFFFFFFFF)
    cs_gpu = int((got[live, 0] ^ (got[live, 1] << np.uint64(1))).sum(dtype=np.uint64))
    ok = bool(hits == int(live.sum()) and cs_cpu == cs_gpu)
    # device-resident: each GPU resolves its share from device arrays, all GPUs at once; device time = max over GPUs
    share = [(nq * k // G, nq * (k + 1) // G) for k in range(G)]
    bufs = []
    for k, (q0, q1) in enumerate(share):
        dv = torch.device("cuda", devs[k])
        bufs.append((torch.from_numpy(rec[q0:q1].view(np.int32)).to(dv), torch.from_numpy(fld[q0:q1].view(np.int32)).to(dv),
                     torch.empty((q1 - q0, 2), dtype=torch.int64, device=dv)))
    streams = [torch.cuda.ExternalStream(m.stream(k), device=torch.device("cuda", devs[k])) for k in range(G)]

    def launch_all(reps):
        evs = []
        for k in range(G):
            with torch.cuda.device(devs[k]):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(streams[k])
                for _ in range(reps):
                    mi.seek_fields_device(k, bufs[k][0].data_ptr(), bufs[k][1].data_ptr(), bufs[k][0].numel(), bufs[k][2].data_ptr())
                e1.record(streams[k])
                evs.append((e0, e1))
        for k in range(G):
            with torch.cuda.device(devs[k]):
                torch.cuda.synchronize()
        return max(e0.elapsed_time(e1) for e0, e1 in evs) / reps
    if len(set(devs)) == G:
        launch_all(3)
        ms = launch_all(20)
    else:
        ms = float("nan")
    segs = mi.segments()
    remote = 1.0 - 1.0 / G
    print(json.dumps({"bench": "multi_seek_fields", "gpus": G, "workload": "cfg5: 256-field CSV, index distributed over the GPUs",
                      "csv_bytes": n, "index_entries": int(len(mi)), "segments": segs, "queries": nq, "index_build_s": build_s,
                      "device_resident": {"ms_per_batch": ms, "value": nq / (ms * 1e-3) / 1e6, "unit": "Mqueries/s",
                                          "remote_fraction": remote,
                                          "nvlink_read_gbs_per_gpu": 16 * (nq / G) * remote / (ms * 1e-3) / 1e9,
                                          "note": "every GPU resolves nq / G queries at once; a slot owned by another GPU is an "
                                                  "8-byte load through the peer mapping (16 B per remote query over NVLink)"},
                      "host_arrays_pinned": {"seconds": best, "value": nq / best / 1e6, "unit": "Mqueries/s"},
                      "parity": {"checked": True, "ok": ok, "what": "hit count and checksum of all (start, end) pairs == oracle seek_field"}}),
          flush=True)
    assert ok
    mi.free()
    m.close()


if __name__ == "__main__":
    main()
