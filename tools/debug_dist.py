"""2-GPU debug: sharded build vs the oracle on a small cfg4 workload (run under torchrun)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
import csv_simd_b200 as cs
from csv_simd_b200 import dist as csd
from bench import make_workload
from oracle import oracle as O

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
size = int(sys.argv[1]) if len(sys.argv) > 1 else (8 << 20)
data, _ = make_workload("cfg4_sharded", rank, world, size)
n = data.size
t = torch.tensor([n], dtype=torch.int64, device=dev); allsz = torch.empty(world, dtype=torch.int64, device=dev)
dist.all_gather_into_tensor(allsz, t); sizes = allsz.cpu().tolist(); goff = sum(sizes[:rank])
ctx = cs.Context(lr); ctx.set_stream(torch.cuda.current_stream(dev).cuda_stream)
d_in = torch.empty(n + 64, dtype=torch.uint8, device=dev); d_in[:n].copy_(torch.from_numpy(data)); torch.cuda.synchronize()
p_host = ctx.shard_quote_parity(d_in.data_ptr(), n)
p_or, c0, s = O.shard_summary(data)
sh = csd.sharded_index_build(ctx, d_in.data_ptr(), n, goff)
got = sh.local.to_host()
want, _ = O.read_closed_form(data, sh.carry_in, goff, with_sentinel=(rank == 0))
idx2 = ctx.index_build_shard_device(d_in.data_ptr(), n, sh.carry_in, goff, rank == 0)
print(f"rank {rank}: n={n} goff={goff} parity dev={p_host} oracle={p_or} ps={sh.parities} carry={sh.carry_in} "
      f"len={got.size} want={want.size} equal={got.size == want.size and bool((got == want).all())} "
      f"explicit_len={len(idx2)} base={sh.base} total={sh.total_len} (c0={c0}, s={s})", flush=True)
dist.destroy_process_group()
