"""Summarise an .ncu-rep: headline metrics + SASS-level stall/instruction hot spots grouped by barrier regions.
usage: python tools/ncu_summary.py report.ncu-rep [--top N]"""
import csv
import subprocess
import sys


def num(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return 0.0


def main():
    rep = sys.argv[1]
    top_n = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 25
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, r = rows[0], rows[1], rows[2]
    keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
            "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
            "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
            "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
            "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "lts__t_sector_hit_rate.pct",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
    for i, h in enumerate(hdr):
        if h in keys or h.startswith("smsp__average_warps_issue_stalled") and "per_issue_active" in h:
            v = r[i]
            if h.startswith("smsp__average_warps_issue_stalled") and num(v) < 0.3:
                continue
            print(f"{h:85s} {units[i]:12s} {v}")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data, seen = [], set()
    for r in rows[2:]:
        if len(r) != len(hdr) or r[0] == "Address":
            continue
        if r[0] in seen:
            break
        seen.add(r[0])
        data.append(r)
    ts = sum(num(r[ix["# Samples"]]) for r in data)
    ti = sum(num(r[ix["Instructions Executed"]]) for r in data)
    print(f"\nSASS instrs {len(data)}, samples {ts:.0f}, warp-inst executed {ti / 1e6:.1f}M")
    reg, cur = [], dict(start=0, s=0, i=0, n=0, ops={})
    for k, r in enumerate(data):
        s_, i_ = num(r[ix["# Samples"]]), num(r[ix["Instructions Executed"]])
        toks = r[ix["Source"]].split()
        op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
        cur["s"] += s_
        cur["i"] += i_
        cur["n"] += 1
        cur["ops"][op] = cur["ops"].get(op, 0) + i_
        if "BAR.SYNC" in r[ix["Source"]] or "SYNCS.PHASECHK" in r[ix["Source"]]:
            cur["end"] = k
            reg.append(cur)
            cur = dict(start=k + 1, s=0, i=0, n=0, ops={})
    cur["end"] = len(data) - 1
    reg.append(cur)
    for g in reg:
        if g["i"] < ti * 0.002 and g["s"] < ts * 0.005:
            continue
        ops = " ".join(f"{a}:{b / 1e6:.1f}" for a, b in sorted(g["ops"].items(), key=lambda x: -x[1])[:8])
        print(f"sass[{g['start']:4d}-{g['end']:4d}] samples {100 * g['s'] / ts:5.1f}%  inst {100 * g['i'] / ti:5.1f}% "
              f"({g['i'] / 1e6:6.1f}M)  {ops}")
    print("\ntop stalled instructions (samples, executed, sass):")
    for k, r in sorted(enumerate(data), key=lambda kr: -num(kr[1][ix["# Samples"]]))[:top_n]:
        print(f"{k:5d} {r[ix['# Samples']]:>7s} {r[ix['Instructions Executed']]:>10s}  {r[ix['Source']][:110]}")


if __name__ == "__main__":
    main()
