"""Multi-GPU sharded index build: contiguous byte ranges, one process per GPU.

Implements the reference README's undone idea "splitting work without first knowing record breaks
(requires toggling interpretation if/when start in quoted text)" (README.md:24):

  pass A   every rank: quote parity p_k of its own shard          (csvb200_shard_quote_parity)
  exchange ONE tiny all_gather of the p_k (NCCL over NVLink; gloo in the CPU tests)
           carry-in parity of rank k = XOR of p_j for j < k
  pass B   every rank: fused index build with that carry-in parity and its global byte offset
           (csvb200_index_build_shard_device); the sentinel entry is emitted by rank 0 only
  exchange a second tiny all_gather of the per-rank entry counts -> global index base of each
           segment (needed for record numbering / lookup routing, not for pass B)

The index stays distributed: rank k holds entries [base_k, base_k + len_k) of the global index.
No bulk data ever crosses GPUs.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import torch
import torch.distributed as dist

from . import api


def carry_in_parities(parities: Sequence[int]) -> List[int]:
    """Exclusive XOR-scan: parity entering shard k given each shard's own quote parity."""
    out, acc = [], 0
    for p in parities:
        out.append(acc)
        acc ^= int(p) & 1
    return out


def exclusive_bases(counts: Sequence[int]) -> List[int]:
    out, acc = [], 0
    for c in counts:
        out.append(acc)
        acc += int(c)
    return out


def _all_gather_i64(value: int, group=None, device: Optional[torch.device] = None) -> List[int]:
    world = dist.get_world_size(group)
    t = torch.tensor([int(value)], dtype=torch.int64, device=device)
    out = torch.empty(world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(out, t, group=group)
    return [int(v) for v in out.cpu().tolist()]


def exchange_parity(local_parity: int, group=None, device: Optional[torch.device] = None):
    """all_gather the shard parities; returns (carry-in parity of this rank, all parities)."""
    ps = _all_gather_i64(local_parity & 1, group, device)
    return carry_in_parities(ps)[dist.get_rank(group)], ps


def exchange_counts(local_count: int, group=None, device: Optional[torch.device] = None):
    """all_gather the per-rank entry counts; returns (global base of this rank's segment, total)."""
    cs = _all_gather_i64(local_count, group, device)
    return exclusive_bases(cs)[dist.get_rank(group)], sum(cs)


@dataclass
class ShardedIndex:
    local: api.StructureIndex   # this rank's segment (global byte positions)
    base: int                   # global slot of local[0]
    total_len: int              # length of the whole (distributed) index, sentinel included
    carry_in: int               # quote parity entering this shard
    parities: List[int]


def sharded_index_build(ctx: api.Context, dev_ptr: int, n: int, global_offset: int, group=None,
                        resolve: bool = True) -> ShardedIndex:
    """Index this rank's shard [global_offset, global_offset + n) of a file split across the ranks
    of `group` at arbitrary byte offsets.

    Everything between pass A and the end of pass B is stream-ordered on the current CUDA stream
    (which the Context must be bound to, see Context.set_stream): parity kernel -> NCCL all_gather
    of 4 bytes per rank -> build kernel that XORs the gathered parities of the lower ranks on the
    device -> NCCL all_gather of the entry counts.  The host only synchronises once, at the end,
    to learn the segment length (resolve=False skips even that; fields base/total_len/carry_in are
    then -1 and `counts` holds the device tensor)."""
    rank = dist.get_rank(group)
    world = dist.get_world_size(group)
    device = torch.device("cuda", ctx.device)
    par_local = torch.empty(1, dtype=torch.int32, device=device)
    ctx.shard_quote_parity_device(dev_ptr, n, par_local.data_ptr())              # pass A
    pars = torch.empty(world, dtype=torch.int32, device=device)
    dist.all_gather_into_tensor(pars, par_local, group=group)                   # 4 bytes per rank over NVLink
    res_local = torch.empty(2, dtype=torch.int64, device=device)
    idx = ctx.index_build_shard_device_ex(dev_ptr, n, pars.data_ptr(), rank, global_offset,
                                          emit_sentinel=(rank == 0), d_result_out=res_local.data_ptr())  # pass B
    res_all = torch.empty(2 * world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(res_all, res_local, group=group)                # entry counts / end parities
    idx._keepalive = (par_local, pars, res_local, res_all)
    if not resolve:
        out = ShardedIndex(idx, -1, -1, -1, [])
        out.counts = res_all
        return out
    host = res_all.cpu().tolist()                                               # the only host sync
    counts = [int(c) for c in host[0::2]]
    counts[0] += 1                                                              # sentinel lives on rank 0
    ps = [int(v) for v in pars.cpu().tolist()]
    return ShardedIndex(idx, exclusive_bases(counts)[rank], sum(counts), carry_in_parities(ps)[rank], ps)
