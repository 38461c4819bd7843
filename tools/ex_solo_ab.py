"""One GPU: the plain build against the exchange-form build (world 1: nothing to wait for) on the same 1 GiB quote-heavy
bytes, interleaved; kernel time from CUDA events around each call on the context's stream."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import csv_simd_b200 as cs
from tools import gen

dev = torch.device("cuda", 0)
data, _ = gen.quoted(1 << 30, seed=44)
n = int(data.size)
d = torch.empty(n + 64, dtype=torch.uint8, device=dev)
d[:n].copy_(torch.from_numpy(data))
ctx = cs.Context(0)
stream = torch.cuda.Stream(dev)
ctx.set_stream(stream.cuda_stream)
ex = ctx.exchange(0, 1)
cs.Exchange.connect_local([ex])
torch.cuda.synchronize()


def timed(fn, reps=1):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    idx = fn()
    e1.record(stream)
    e1.synchronize()
    idx.sync()
    idx.free()
    return e0.elapsed_time(e1)


plain = lambda: ctx.index_build_device(d.data_ptr(), n)
exch = lambda: ctx.index_build_shard_exchange(ex, d.data_ptr(), n, 0)
for _ in range(3):
    timed(plain); timed(exch)
tp, te = [], []
for _ in range(30):
    tp.append(timed(plain)); te.append(timed(exch))
tp.sort(); te.sort()
print(json.dumps({"plain_ms_med": tp[15], "plain_ms_min": tp[0], "exchange_form_ms_med": te[15], "exchange_form_ms_min": te[0],
                  "note": "world 1: rank 0 posts to itself, waits for nobody; includes the conditional re-index launch that exits at once"}))
