"""Multi-GPU check of the whole chain on a sharded file (run under torchrun, one rank per GPU):
sharded speculative build -> gather_segments / replicate_index -> batched lookups on EVERY rank, compared with
the oracle's index and seek_field over the full file.  Prints one JSON line on rank 0.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_lookup_check.py [bytes]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import csv_simd_b200 as cs  # noqa: E402
from csv_simd_b200 import dist as csd  # noqa: E402
from oracle import oracle as O  # noqa: E402
from tools import gen  # noqa: E402


def main():
    size = int(sys.argv[1]) if len(sys.argv) > 1 else (256 << 20)
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    # every rank generates the same full file (fixed-width rows so the tape is valid), then takes its byte range
    data, rows = gen.unquoted(size, seed=61, nfields=32, modulus=10 ** 9)
    n = data.size
    cuts = [0] + [(k * n) // world + 37 * k + 13 for k in range(1, world)] + [n]
    lo, hi = cuts[rank], cuts[rank + 1]
    shard = torch.from_numpy(data[lo:hi].copy()).to(dev)
    ctx = cs.Context(local)
    stream = torch.cuda.current_stream(dev)
    ctx.set_stream(stream.cuda_stream)
    sh = csd.sharded_index_build(ctx, shard.data_ptr(), hi - lo, lo)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    full = csd.replicate_index(ctx, sh, sh.counts, n)
    torch.cuda.synchronize()
    rep_s = time.perf_counter() - t0
    want = O.read_sse(data)
    got = full.to_host()
    ok_index = got.shape == want.shape and bool((got == want).all())
    rc, jump = full.tape_init(32, False)
    rec, fld = gen.queries(2_000_000, rc, 32, seed=62 + rank)        # different queries on every rank
    d_rec, d_fld = torch.from_numpy(rec.view(np.int32)).to(dev), torch.from_numpy(fld.view(np.int32)).to(dev)
    d_out = torch.empty((rec.size, 2), dtype=torch.int64, device=dev)
    full.seek_fields_device(d_rec.data_ptr(), d_fld.data_ptr(), rec.size, d_out.data_ptr())
    torch.cuda.synchronize()
    res = d_out.cpu().numpy().view(np.uint64)
    cs_cpu, hits = O.seek_fields_timed(want, n, rc, 32, False, rec, fld)
    live = res[:, 0] != np.uint64(0xFFFFFFFFFFFFFFFF)
    cs_gpu = int((res[live, 0] ^ (res[live, 1] << np.uint64(1))).sum(dtype=np.uint64))
    ok_seek = hits == int(live.sum()) and cs_cpu == cs_gpu
    flags = torch.tensor([int(ok_index), int(ok_seek)], dtype=torch.int64, device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"check": "sharded build -> replicated index -> lookups on every rank", "world": world,
                          "file_bytes": int(n), "index_entries": int(want.size), "index_equals_oracle_on_all_ranks": bool(flags[0].item()),
                          "lookups_equal_oracle_on_all_ranks": bool(flags[1].item()), "replicate_ms": rep_s * 1e3,
                          "replicate_gbs": 8 * want.size / rep_s / 1e9}), flush=True)
    full.free()
    sh.local.free()
    dist.destroy_process_group()
    ctx.close()
    if not (flags[0].item() and flags[1].item()):
        sys.exit(1)


if __name__ == "__main__":
    main()
