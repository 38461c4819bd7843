#!/usr/bin/env python
"""bench.py -- CSV bytes indexed per second (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

A "step" is one pass of the hot path (csv bytes -> structural index) over one batch of synthetic
CSV.  N=1: BASELINE config 2 (1 GiB unquoted, 16 numeric fields/row, LF).  N>1: BASELINE config 4
(quote-heavy CRLF grammar, N GiB sharded at arbitrary byte offsets, 1 GiB per GPU, weak scaling)
with the cross-shard quote-parity exchange inside every step.

  value      whole-job CSV GB/s with the input already resident in HBM (max over ranks, CUDA events)
  e2e        the same metric through the public host-buffer C-ABI call: pinned host bytes in,
             host index out, H2D + kernels + D2H all inside the timed region
  roofline   dominant kernel (index_build_kernel): algorithmic bytes N + 8E per launch / average
             launch duration (CUDA events around each launch) vs the measured HBM copy peak
  cpu_baseline / --impl reference
             the reference's CPU path (oracle/csv_oracle.c: the literal SSE restatement, 1 thread
             because the reference is a single serial loop) on the box's host cores
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GiB = 1 << 30
METRIC = "csv_bytes_indexed_per_sec"
UNIT = "GB/s"


# ---------------------------------------------------------------------------------------------
# workloads (deterministic generators, tools/gen_csv.c)
# ---------------------------------------------------------------------------------------------
def make_workload(name: str, rank: int, world: int, size: int):
    """Returns (bytes ndarray for this rank, global byte offset of the shard, description)."""
    from tools import gen
    if name == "cfg2_unquoted":
        data, rows = gen.unquoted(size, seed=42, first_row=rank << 32, with_header=(rank == 0))
        desc = "synthetic unquoted CSV, 16 numeric fields/row, LF (BASELINE config 2)"
    elif name == "cfg3_quoted":
        data, rows = gen.quoted(size, seed=43, first_row=rank << 32, with_header=(rank == 0))
        desc = "synthetic quote-heavy CSV: embedded commas, CRLF, newlines and \"\" escapes (BASELINE config 3)"
    elif name == "cfg4_sharded":
        # logical file = concatenation of per-rank row streams; shard k is cut at start(G_k) + 37k + 13,
        # i.e. NOT at a record boundary and not 16-byte aligned in the file: rank k owns
        # G_k[d_k:] ++ G_{k+1}[:d_{k+1}]
        def delta(k):
            return 0 if k == 0 or k >= world else 37 * k + 13
        own, rows = gen.quoted(size, seed=44, first_row=rank << 32, with_header=(rank == 0))
        parts = [own[delta(rank):]]
        if rank + 1 < world:
            nxt, _ = gen.quoted(4096, seed=44, first_row=(rank + 1) << 32, with_header=False)
            parts.append(nxt[:delta(rank + 1)])
        data = np.concatenate(parts)
        desc = ("synthetic quote-heavy CSV sharded at arbitrary byte offsets with cross-shard quote-parity "
                "fix-up (BASELINE config 4)")
    else:
        raise SystemExit(f"unknown workload {name}")
    return np.ascontiguousarray(data), desc


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed regions (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,utilization.gpu,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                util = float(r[4])
                if util < 5 and len(self.rows) > 3:
                    continue
                sm.append(float(r[1]))
                smax.append(float(r[2]))
                power.append(float(r[3]))
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except (ValueError, IndexError):
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(workload: str, n: int):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the hot kernel, from the committed
    ncu --set full capture of the same 1 GiB workload (profiles/); (None, why) when no capture matches.
    cfg4 shards are 1 GiB of the cfg3 grammar (another seed), so the cfg3 capture stands in for them."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_v6_traffic.json")) as f:
            t = json.load(f)
        key = "cfg2" if workload == "cfg2_unquoted" else "cfg3"
        if abs(n - GiB) > (1 << 20):
            return None, "no ncu capture at this size"
        note = t[key]["source"] + ("; cfg4 shard = same grammar and size as cfg3" if workload == "cfg4_sharded" else "")
        return t[key]["traffic_bytes_per_launch"], note
    except Exception as e:  # noqa: BLE001
        return None, f"profiles/r01_v6_traffic.json unreadable: {e}"


def host_cpu():
    model = "unknown"
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                model = line.split(":", 1)[1].strip()
                break
    except OSError:
        pass
    return model, os.cpu_count()


def cpu_reference_gbs(data: np.ndarray, reps: int):
    """The reference's CPU path (literal SSE restatement, single thread) on this host."""
    from oracle import oracle as O
    buf = O.aligned_copy(data)
    buf.sum()  # pre-fault (the reference's mmap page faults are excluded, as BASELINE.md states)
    best, E = 0.0, 0
    for _ in range(reps):
        t = time.perf_counter()
        E, _ = O.read_sse_timed(buf)
        dt = time.perf_counter() - t
        best = max(best, buf.size / dt / 1e9)
    return best, E


# ---------------------------------------------------------------------------------------------
def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    wl = args.workload or ("cfg2_unquoted" if world == 1 else "cfg4_sharded")
    size = args.size or GiB
    data, desc = make_workload(wl if wl != "cfg4_sharded" else "cfg3_quoted", 0, 1, size)
    from oracle import oracle as O
    buf = O.aligned_copy(data)
    buf.sum()
    for _ in range(min(args.warmup, 1)):
        O.read_sse_timed(buf)
    t = time.perf_counter()
    for _ in range(args.steps):
        E, _ = O.read_sse_timed(buf)
    dt = time.perf_counter() - t
    gbs = buf.size * args.steps / dt / 1e9
    model, ncpu = host_cpu()
    sample = (f"{buf.size} bytes ({wl} grammar, one shard) per step, single thread: the reference is one serial "
              f"loop (src/reader.rs:229-258); host {model}, {ncpu} logical cpus")
    line = {
        "impl": "reference", "metric": METRIC, "value": gbs, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": wl, "description": desc, "bytes_per_step": int(buf.size), "entries": int(E)},
        "cpu_baseline": {"value": gbs, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample},
        "e2e": {"value": gbs, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_ours(args, rank: int, local_rank: int, world: int):
    import torch
    import torch.distributed as dist

    import csv_simd_b200 as cs
    from csv_simd_b200 import dist as csd

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    from csv_simd_b200 import numa
    numa_node = numa.bind_to_device(local_rank)   # pinned staging buffers on the GPU's own NUMA node
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    wl = args.workload or ("cfg2_unquoted" if world == 1 else "cfg4_sharded")
    size = args.size or GiB
    data, desc = make_workload(wl, rank, world, size)
    n = int(data.size)
    sizes = [n]
    if world > 1:
        t = torch.tensor([n], dtype=torch.int64, device=dev)
        allsz = torch.empty(world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(allsz, t)
        sizes = [int(v) for v in allsz.cpu().tolist()]
    goff = sum(sizes[:rank])
    total_bytes = sum(sizes)

    ctx = cs.Context(local_rank)
    stream = torch.cuda.current_stream(dev)
    ctx.set_stream(stream.cuda_stream)  # our kernels run on torch's current stream: torch events bracket them

    d_in = torch.empty(n + 64, dtype=torch.uint8, device=dev)
    d_in[:n].copy_(torch.from_numpy(data))
    torch.cuda.synchronize(dev)

    def step_device(resolve=True):
        # resolve=False: nothing in the step waits on the host (the timed loops); the index, its length
        # and (N>1) the all-gathered counts are complete in HBM when the stream reaches the end event
        if world == 1:
            idx = ctx.index_build_device(d_in.data_ptr(), n)
            if resolve:
                idx.sync()
            return idx
        return csd.sharded_index_build(ctx, d_in.data_ptr(), n, goff, resolve=resolve).local

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()

    # ---- warm-up + correctness of the configuration (entry count vs the oracle happens in tests/) ----
    E, carry_in = 0, 0
    for _ in range(max(args.warmup, 3)):
        idx = step_device()
        E = len(idx)
        idx.free()
    if world > 1:
        sh = csd.sharded_index_build(ctx, d_in.data_ptr(), n, goff)
        carry_in, E = sh.carry_in, len(sh.local)
        sh.local.free()

    # ---- value: K device-resident steps, CUDA events, max over ranks ----
    launches0 = ctx.launch_count()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        idx = step_device(resolve=False)
        idx.free()
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = ctx.launch_count() - launches0
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
        t = torch.tensor([launches, E], dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        launches, E_total = int(t[0].item()), int(t[1].item())
    else:
        E_total = E
    ms_per_step = ms_total / args.steps
    value = total_bytes / (ms_per_step * 1e-3) / 1e9

    # ---- roofline of the dominant kernel: per-launch duration from events around each launch ----
    kms = []
    for _ in range(min(args.steps, 50)):
        if world == 1:
            idx = ctx.index_build_device(d_in.data_ptr(), n)
        else:
            idx = ctx.index_build_shard_device(d_in.data_ptr(), n, carry_in, goff, rank == 0)
        idx.sync()
        kms.append(ctx.last_build_ms())
        E_local = len(idx)
        idx.free()
    k_ms = sum(kms) / len(kms)
    alg_bytes = n + 8 * E_local
    peak, peak_src = hbm_peak()
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9

    # ---- e2e: pinned host bytes -> C-ABI -> host index (H2D + kernels + D2H inside the timed region) ----
    h_in = torch.from_numpy(data).pin_memory()
    h_out = torch.empty(E_local + 1024, dtype=torch.int64).pin_memory()
    e2e_steps = max(3, min(args.steps, 10))

    def step_e2e():
        if world == 1:
            return ctx.index_build_to_host(h_in.data_ptr(), n, h_out.data_ptr(), h_out.numel())
        ln, _base, _total, _redone = csd.sharded_index_build_to_host(ctx, h_in.data_ptr(), n, goff, h_out.data_ptr(),
                                                                      h_out.numel())
        return ln

    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ln = step_e2e()
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_gbs = total_bytes / e2e_s / 1e9
    assert ln == E_local, (ln, E_local)

    # the same call on ordinary (pageable) memory, as a caller holding an mmap and a Vec would make it: informational
    e2e_pageable = None
    if world == 1:
        out_pg = np.empty(E_local + 1024, dtype=np.uint64)
        out_pg[::512] = 0
        ctx.index_build_to_host(data.ctypes.data, n, out_pg.ctypes.data, out_pg.size)
        t0 = time.perf_counter()
        for _ in range(3):
            ctx.index_build_to_host(data.ctypes.data, n, out_pg.ctypes.data, out_pg.size)
        e2e_pageable = n / ((time.perf_counter() - t0) / 3) / 1e9

    clocks = sampler.stop() if rank == 0 else None

    # ---- CPU baseline (rank 0, N=1 only): the reference's CPU path on this box's host cores ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        gbs, E_cpu = cpu_reference_gbs(data, reps=5)
        model, ncpu = host_cpu()
        assert E_cpu == E_local, (E_cpu, E_local)  # the oracle checks the GPU's entry count
        cpu = {"value": gbs, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": (f"full {n}-byte input, best of 5 passes, single thread (the reference is one serial loop, "
                          f"src/reader.rs:229-258); input pre-faulted in RAM, println!s omitted; host {model}, "
                          f"{ncpu} logical cpus")}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": wl, "description": desc, "bytes_per_gpu": n, "total_bytes": total_bytes,
                       "index_entries": E_total, "index_entry_bytes": 8,
                       "l2_policy": "input (1 GiB) and index (>= 0.5 GB) are far larger than the 126 MB L2; no flush",
                       "parallelism": f"byte-range shards x{world}" if world > 1 else "single GPU",
                       "host_numa_node_rank0": numa_node},
            "roofline": {"bound": "hbm", "kernel": "index_build_tma_kernel", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic(wl, n)[0], "traffic_source": ncu_traffic(wl, n)[1],
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": k_ms,
                         "csv_gbs_kernel_only": n / (k_ms * 1e-3) / 1e9},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_gbs, "unit": UNIT, "h2d_bytes_per_step": n, "d2h_bytes_per_step": 8 * E_local,
                    "ms_per_step": e2e_s * 1e3, "steps": e2e_steps, "pageable_buffers_gbs": e2e_pageable,
                    "api": "csvb200_index_build_to_host" if world == 1 else
                           "csvb200_shard_build_to_host + all_gather + csvb200_shard_job_verify (csv_simd_b200.dist.sharded_index_build_to_host)"},
            "gpu_launches": launches,
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    ctx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=[None, "cfg2_unquoted", "cfg3_quoted", "cfg4_sharded"])
    ap.add_argument("--size", type=int, default=None, help="bytes per GPU (default 1 GiB)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        if args.steps > 20:
            args.steps = 20  # 1 GiB per step at ~1 GB/s: keep the arm within a couple of minutes
        run_reference(args, rank, world)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("launch with torchrun: python -m torch.distributed.run --nproc-per-node N bench.py --gpus N")
    run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
