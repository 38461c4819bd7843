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


def sharded_index_build(ctx: api.Context, dev_ptr: int, n: int, global_offset: int, group=None) -> ShardedIndex:
    """Index this rank's shard [global_offset, global_offset + n) of a file split across the ranks
    of `group` at arbitrary byte offsets."""
    rank = dist.get_rank(group)
    device = torch.device("cuda", ctx.device)
    p = ctx.shard_quote_parity(dev_ptr, n)                       # pass A
    carry, ps = exchange_parity(p, group, device)                # 8 bytes per rank over NVLink
    idx = ctx.index_build_shard_device(dev_ptr, n, carry, global_offset, emit_sentinel=(rank == 0))  # pass B
    base, total = exchange_counts(len(idx), group, device)
    return ShardedIndex(idx, base, total, carry, ps)
