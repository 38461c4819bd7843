"""Multi-GPU sharded index build: contiguous byte ranges, one process per GPU.

Implements the reference README's undone idea "splitting work without first knowing record breaks
(requires toggling interpretation if/when start in quoted text)" (README.md:24):

two-pass protocol:
  pass A   every rank: quote parity p_k of its own shard          (csvb200_shard_quote_parity)
  exchange ONE tiny all_gather of the p_k (NCCL over NVLink; gloo in the CPU tests)
           carry-in parity of rank k = XOR of p_j for j < k
  pass B   every rank: fused index build with that carry-in parity and its global byte offset
           (csvb200_index_build_shard_device); the sentinel entry is emitted by rank 0 only
  exchange a second tiny all_gather of the per-rank entry counts -> global index base of each
           segment (needed for record numbering / lookup routing, not for pass B)

speculative protocol (default; one pass over the bytes and one exchange in the common case):
  predict  every rank guesses its carry-in parity from the first unambiguous quote of its shard
  build    fused index build with the guess; reports {entries, end parity, carry used, separator total}
  exchange ONE all_gather of those 32 bytes per rank
  verify   every rank: true carries = exclusive XOR-scan of (end ^ used); true counts from
           c0 + c1 = total; a shard whose guess was wrong (and only that shard) is re-indexed

exchange protocol (round 2; pass exchange=make_exchange(ctx)): the same speculative scheme with the collective
replaced by peer-mapped mailboxes -- every rank's 32-byte row is written straight into every peer's HBM over NVLink
from inside the index-build launch, each rank resolves the carries of the LOWER ranks on the device, and the step is
predictor + build launch + conditional re-index launch: no NCCL call, no verify launch, no host round trip
(csvb200_index_build_shard_exchange).  torch.distributed is only used once, to hand the 64-byte IPC handles around.

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


def verify_speculation(gathered: Sequence[Sequence[int]]):
    """Host statement of what csvb200_index_shard_verify computes on the device (verify_carry_kernel):
    gathered[k] = (entries emitted under the carry rank k used, end parity under that carry, carry used,
    separator total of shard k).  Returns (true carries, true entry counts, redo flags).  A shard's
    own quote parity is end ^ used whatever carry it guessed, and flipping its carry swaps the separators
    inside and outside quotes, so its true count is `entries` or `total - entries`."""
    carries, counts, redo, carry = [], [], [], 0
    for cnt, endp, used, total in gathered:
        carries.append(carry)
        counts.append(int(cnt) if carry == (int(used) & 1) else int(total) - int(cnt))
        redo.append(carry != (int(used) & 1))
        carry ^= (int(endp) ^ int(used)) & 1
    return carries, counts, redo


def make_exchange(ctx: api.Context, group=None) -> api.Exchange:
    """This rank's endpoint of the mailbox exchange, connected to every rank of `group` (processes of one node):
    one all_gather of the 64-byte CUDA IPC handles at set-up time, nothing collective afterwards."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    ex = ctx.exchange(rank, world)
    if world > 1:
        handles = [None] * world
        dist.all_gather_object(handles, ex.handle(), group=group)
        ex.connect(handles)
        dist.barrier(group=group)     # every rank has mapped every mailbox before the first build posts
    return ex


@dataclass
class ShardedIndex:
    local: api.StructureIndex   # this rank's segment (global byte positions)
    base: int                   # global slot of local[0]
    total_len: int              # length of the whole (distributed) index, sentinel included
    carry_in: int               # quote parity entering this shard
    parities: List[int]
    counts: object = None       # resolved: entries per rank (rank 0 incl. the sentinel); resolve=False: the device tensor


def sharded_index_build(ctx: api.Context, dev_ptr: int, n: int, global_offset: int, group=None,
                        resolve: bool = True, speculative: bool = True, predict_window: int = 0,
                        exchange: Optional[api.Exchange] = None) -> ShardedIndex:
    """Index this rank's shard [global_offset, global_offset + n) of a file split across the ranks
    of `group` at arbitrary byte offsets.  Everything is stream-ordered on the current CUDA stream
    (which the Context must be bound to, see Context.set_stream); the host only synchronises once, at
    the end, to learn the segment length (resolve=False skips even that; fields base/total_len/carry_in
    are then -1 and `counts` holds the device tensor of per-rank {entries, carry}).

    speculative=True (default): ONE exchange.  Every rank predicts its carry-in parity from the first
    unambiguous quote of its shard, indexes at once, then a single NCCL all_gather of 32 bytes per rank
    ({entries, end parity, carry used, separator total}) lets every rank verify the carry chain on the
    device, derive every shard's true entry count (c0 + c1 = total) and re-index only a mispredicted
    shard (csvb200_index_build_shard_speculative / csvb200_index_shard_verify).

    speculative=False: the two-pass protocol of SURVEY.md 8e -- parity kernel -> all_gather of 4 bytes
    per rank -> build kernel that XORs the gathered parities of the lower ranks on the device ->
    all_gather of the entry counts."""
    rank = dist.get_rank(group)
    world = dist.get_world_size(group)
    device = torch.device("cuda", ctx.device)
    if exchange is not None:
        # exchange=...: the whole step is ONE C call; the rows cross NVLink from inside the build launch
        idx = ctx.index_build_shard_exchange(exchange, dev_ptr, n, global_offset, predict_window)
        if not resolve:
            return ShardedIndex(idx, -1, -1, -1, [])
        info = idx.shard_info()                              # the only host sync: this rank's launches
        counts, carries = exchange.counts(idx)               # host-side wait for the rows of ALL ranks
        ps = [carries[k] ^ (carries[k + 1] if k + 1 < world else 0) for k in range(world)]   # informational
        return ShardedIndex(idx, info["base"], sum(counts), info["carry_in"], ps, counts)
    if speculative:
        res_local = torch.empty(4, dtype=torch.int64, device=device)
        idx = ctx.index_build_shard_speculative(dev_ptr, n, rank, global_offset, emit_sentinel=(rank == 0),
                                                d_result_out=res_local.data_ptr(), predict_window=predict_window)
        res_all = torch.empty(4 * world, dtype=torch.int64, device=device)
        dist.all_gather_into_tensor(res_all, res_local, group=group)            # the only exchange: 32 B per rank
        final = torch.empty(2 * world, dtype=torch.int64, device=device)
        idx.shard_verify(res_all.data_ptr(), world, final.data_ptr())           # + conditional re-index
        idx._keepalive = (res_local, res_all, final)
        if not resolve:
            out = ShardedIndex(idx, -1, -1, -1, [])
            out.counts = final
            return out
        host = torch.cat([final, res_all]).cpu().tolist()                       # the only host sync
        counts = [int(c) for c in host[0:2 * world:2]]
        counts[0] += 1                                                          # sentinel lives on rank 0
        carries = [int(c) for c in host[1:2 * world:2]]
        g = host[2 * world:]
        ps = [(int(g[4 * k + 1]) ^ int(g[4 * k + 2])) & 1 for k in range(world)]  # shard parity = end ^ carry used
        return ShardedIndex(idx, exclusive_bases(counts)[rank], sum(counts), carries[rank], ps, counts)
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
    return ShardedIndex(idx, exclusive_bases(counts)[rank], sum(counts), carry_in_parities(ps)[rank], ps, counts)


def sharded_index_build_to_host(ctx: api.Context, host_ptr: int, n: int, global_offset: int, dst_ptr: int, dst_cap: int,
                                group=None, exchange: Optional[api.Exchange] = None):
    """End-to-end form: this rank's shard in (pinned) host memory -> this rank's index segment in host memory
    (csvb200_shard_build_to_host: chunked H2D, chained launches, overlapped D2H under the predicted carry), then
    the one all_gather and csvb200_shard_job_verify.  Returns (entries in dst, base slot of the segment in the
    global index, total length of the global index, whether this shard had to be re-indexed)."""
    rank = dist.get_rank(group)
    world = dist.get_world_size(group)
    device = torch.device("cuda", ctx.device)
    if exchange is not None:
        ln, info, counts = ctx.shard_build_to_host_exchange(exchange, host_ptr, n, global_offset, dst_ptr, dst_cap,
                                                            want_counts=True)
        return ln, info["base"], sum(counts), bool(info["redone"])
    res_local = torch.empty(4, dtype=torch.int64, device=device)
    _, job = ctx.shard_build_to_host(host_ptr, n, rank, global_offset, rank == 0, dst_ptr, dst_cap, res_local.data_ptr())
    res_all = torch.empty(4 * world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(res_all, res_local, group=group)
    final = torch.empty(2 * world, dtype=torch.int64, device=device)
    ln, redone = ctx.shard_job_verify(job, res_all.data_ptr(), world, final.data_ptr())
    counts = [int(c) for c in final.cpu().tolist()[0::2]]
    counts[0] += 1
    return ln, exclusive_bases(counts)[rank], sum(counts), redone


def segment_layout(counts: Sequence[int]):
    """(bases, total) of the per-rank index segments; counts[0] includes the sentinel."""
    bases = exclusive_bases(counts)
    return bases, int(sum(int(c) for c in counts))


def gather_segments(local: torch.Tensor, counts: Sequence[int], group=None) -> torch.Tensor:
    """All ranks' index segments concatenated in rank order, on every rank.  Every segment is broadcast by its owner
    straight into its final place in the result (one collective per rank, NVLink-rate on NCCL): nothing is padded to
    the longest segment and nothing is compacted afterwards.  `local` holds this rank's counts[rank] entries (int64
    view of the u64 positions); works on any backend (NCCL on the GPUs, gloo in the CPU tests)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    counts = [int(c) for c in counts]
    bases, total = segment_layout(counts)
    full = torch.empty(max(total, 1), dtype=torch.int64, device=local.device)
    if counts[rank]:
        full[bases[rank]:bases[rank] + counts[rank]].copy_(local[:counts[rank]])
    works = []
    for k in range(world):
        if counts[k]:
            src = dist.get_global_rank(group, k) if group is not None else k
            works.append(dist.broadcast(full[bases[k]:bases[k] + counts[k]], src=src, group=group, async_op=True))
    for w in works:
        w.wait()
    return full[:total] if total else full[:0]


def replicate_index(ctx: api.Context, sharded: ShardedIndex, counts: Sequence[int], input_bytes: int, group=None):
    """Every rank ends up with the WHOLE index ("segments are all-gathered if a replicated index is
    wanted") as an index object the batched lookups (K4) run on locally: gather_segments + csvb200_index_wrap_device.
    counts = entries per rank, rank 0 including the sentinel (ShardedIndex.counts).  Returns the StructureIndex; the
    gathered tensor that owns its memory rides along as ._keepalive."""
    device = torch.device("cuda", ctx.device)
    rank = dist.get_rank(group)
    seg = _segment_tensor(sharded.local, int(counts[rank]), device)
    full = gather_segments(seg, counts, group)
    idx = ctx.index_wrap_device(full.data_ptr(), int(full.numel()), input_bytes)
    idx._keepalive = full
    return idx


def _segment_tensor(local_index, n_local: int, device) -> torch.Tensor:
    """This rank's segment as a torch view of the library's device memory (zero copy)."""
    class _Raw:
        __cuda_array_interface__ = {"shape": (max(n_local, 1),), "typestr": "<i8",
                                    "data": (local_index.device_ptr, False), "version": 2}
    if n_local == 0:
        return torch.zeros(0, dtype=torch.int64, device=device)
    return torch.as_tensor(_Raw(), device=device)[:n_local]
