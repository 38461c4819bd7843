"""Per-kernel evidence table from an .ncu-rep (ncu --set full): duration, DRAM bytes and achieved DRAM GB/s,
sectors per request (global loads / stores), occupancy, issue utilisation and the top warp-stall reasons.
usage: python tools/ncu_table.py report.ncu-rep [--peak GBS] > profiles/xxx.md"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def num(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return 0.0


UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0,
        "usecond": 1e-6, "nsecond": 1e-9, "msecond": 1e-3, "second": 1.0}


def main():
    rep = sys.argv[1]
    peak = 6453.4
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    if "--peak" in sys.argv:
        peak = float(sys.argv[sys.argv.index("--peak") + 1])
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}

    def val(r, name):
        i = ix.get(name)
        if i is None:
            return 0.0
        return num(r[i]) * UNIT.get(units[i], 1.0)

    stall_cols = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
    print(f"# ncu --set full, one row per launch ({os.path.basename(rep)}); DRAM peak for the fraction = {peak:.1f} GB/s "
          f"(measured copy rate)\n")
    print("| kernel | grid x block | regs | time us | DRAM rd MB | DRAM wr MB | DRAM GB/s | of peak | sect/req ld | sect/req st "
          "| warps active % | issue active % | top stalls (warps per issue) |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
    for r in rows[2:]:
        if len(r) != len(hdr):
            continue
        name = r[ix["Kernel Name"]].split("(")[0].split("::")[-1]
        t = val(r, "gpu__time_duration.sum")
        rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
        gbs = (rd + wr) / t / 1e9 if t else 0.0
        lreq, lsec = val(r, "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum"), val(r, "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum")
        sreq, ssec = val(r, "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum"), val(r, "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum")
        stalls = sorted(((num(r[ix[c]]), c.split("issue_stalled_")[1].split("_per_issue")[0]) for c in stall_cols), reverse=True)[:3]
        st = ", ".join(f"{n_} {v:.2f}" for v, n_ in stalls)
        print(f"| {name} | {r[ix['launch__grid_size']]} x {r[ix['launch__block_size']]} | {r[ix['launch__registers_per_thread']]} "
              f"| {t * 1e6:.1f} | {rd / 1e6:.1f} | {wr / 1e6:.1f} | {gbs:.0f} | {gbs / peak:.2f} "
              f"| {lsec / lreq if lreq else 0:.1f} | {ssec / sreq if sreq else 0:.1f} "
              f"| {num(r[ix['sm__warps_active.avg.pct_of_peak_sustained_active']]):.0f} "
              f"| {num(r[ix['smsp__issue_active.avg.pct_of_peak_sustained_active']]):.0f} | {st} |")


if __name__ == "__main__":
    main()
