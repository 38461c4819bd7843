"""Where the prefix chain's time goes, per super-tile (debug build path: CSVB200_DBG_TIMELINE=1 selects an instrumented
instantiation of index_build_tma_kernel that stamps %clock64 / %globaltimer at five points of every super-tile):

    T_ticket  the producer drew the tile's ticket
    T_agg     the look-back warp has all the tile's warp aggregates (the tile is classified)
    T_pref    its look-back is done (prefix known, published)
    T_need    the workers reach the compaction of the tile (one classification later: kSkew = 1)
    T_got     ... and have its prefix

All five are taken on the tile's own SM (same clock).  Prints one JSON line with the distributions, in microseconds at the
SM clock reported by torch / nvidia-smi, and who the waits follow: the look-back (T_pref - T_agg) or the aggregate itself.

    CSVB200_DBG_TIMELINE=1 python tools/timeline.py [cfg3_quoted|cfg2_unquoted] [bytes]
"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import csv_simd_b200 as cs  # noqa: E402
from tools import gen  # noqa: E402


def pct(a, qs=(50, 90, 99, 100)):
    return {f"p{q}": round(float(np.percentile(a, q)), 3) for q in qs} | {"mean": round(float(a.mean()), 3)}


def main():
    assert os.environ.get("CSVB200_DBG_TIMELINE"), "set CSVB200_DBG_TIMELINE=1"
    wl = sys.argv[1] if len(sys.argv) > 1 else "cfg3_quoted"
    size = int(sys.argv[2]) if len(sys.argv) > 2 else (1 << 30)
    dev = torch.device("cuda", 0)
    ctx = cs.Context(0)
    data, _ = gen.unquoted(size, seed=42) if wl == "cfg2_unquoted" else gen.quoted(size, seed=43)
    n = int(data.size)
    d = torch.empty(n + 64, dtype=torch.uint8, device=dev)
    d[:n].copy_(torch.from_numpy(data))
    torch.cuda.synchronize()
    for _ in range(4):      # warm-up, then the build whose timeline is read
        idx = ctx.index_build_device(d.data_ptr(), n)
        idx.sync()
        ms = ctx.last_build_ms()
        idx.free()
    lib = cs._lib.load()
    lib.csvb200_debug_timeline.restype = C.c_int
    lib.csvb200_debug_timeline.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    words = C.c_size_t()
    lib.csvb200_debug_timeline(ctx._h, None, 0, C.byref(words))
    buf = np.zeros(words.value, dtype=np.uint64)
    rc = lib.csvb200_debug_timeline(ctx._h, buf.ctypes.data, buf.size, C.byref(words))
    assert rc == 0, rc
    tiles = (n + 65535) // 65536
    t = buf[:tiles * 8].reshape(tiles, 8).astype(np.int64)
    mhz = 1965.0
    us = lambda c: c / mhz   # noqa: E731
    t_agg, t_pref, t_need, t_got, meta, g_agg, t_ticket, g_pref = (t[:, k] for k in range(8))
    trips, polls = (t_ticket >> 48) & 0xff, (t_ticket >> 56) & 0xff
    t_ticket = t_ticket & ((1 << 48) - 1)
    t_agg48, t_pref48 = t_agg & ((1 << 48) - 1), t_pref & ((1 << 48) - 1)
    it = (meta >> 32).astype(np.int64)
    cta = (meta >> 16) & 0xffff
    ok = (t_need > 0) & (t_got > 0) & (t_agg > 0)
    wait = us((t_got - t_need)[ok])                 # workers stalled on the prefix
    look = us((t_pref - t_agg)[ok])                 # look-back duration
    slack = us((t_need - t_agg)[ok])                # time between "tile classified" and "prefix needed"
    classify = us((t_agg48 - t_ticket)[ok])           # ticket -> aggregates in (TMA latency + classification)
    late = us((t_pref - t_need)[ok])                # > 0: the prefix came after it was needed
    period = []
    for c in np.unique(cta[ok])[:64]:
        a = np.sort(t_agg[ok & (cta == c)])
        if a.size > 2:
            period.append(np.diff(a).mean())
    total_us = ms * 1e3
    per_cta_wait = np.zeros(int(cta.max()) + 1)
    np.add.at(per_cta_wait, cta[ok], us((t_got - t_need)[ok]))
    # whom the late prefixes follow: wall-clock (globaltimer) gap between this tile's aggregate and the LAST aggregate among
    # the 64 tiles before it
    g = g_agg.astype(np.int64)
    prev_last = np.array([g[max(0, k - 64):k].max() if k else g[0] for k in range(tiles)])
    behind = (prev_last - g)[ok] / 1e3              # us: > 0 = an earlier tile's aggregate came after mine
    out = {"workload": wl, "bytes": n, "tiles": tiles, "kernel_ms_instrumented": ms, "sm_mhz_assumed": mhz,
           "tile_period_us_per_cta": round(float(us(np.mean(period))), 3),
           "ticket_to_aggregates_us": pct(classify), "aggregates_to_need_us (slack)": pct(slack),
           "lookback_us": pct(look), "lookback_round_trips": pct(trips[ok].astype(np.float64)), "lookback_retries": pct(polls[ok].astype(np.float64)),
           "lookback_us_per_round_trip": round(float(look.sum() / max(trips[ok].sum(), 1)), 3), "prefix_after_need_us": pct(np.maximum(late, 0)),
           "workers_wait_us": pct(wait), "share_of_tiles_that_wait_over_0.2us": round(float((wait > 0.2).mean()), 4),
           "wait_us_per_cta_total": pct(per_cta_wait[per_cta_wait > 0]), "kernel_us": round(total_us, 1),
           "wait_share_of_kernel": round(float(per_cta_wait[per_cta_wait > 0].mean() / total_us), 4),
           "latest_aggregate_among_64_predecessors_minus_own_us": pct(behind),
           "first_tiles_of_a_cta (it < 2) share of all wait": round(float(us((t_got - t_need)[ok & (it < 2)]).sum() / max(wait.sum(), 1e-9)), 4)}
    print(json.dumps(out))
    ctx.close()


if __name__ == "__main__":
    main()
