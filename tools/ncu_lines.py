"""Per-CUDA-source-line instruction / stall-sample totals of an .ncu-rep captured with --import-source on.
usage: python tools/ncu_lines.py report.ncu-rep [--top N]"""
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
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 50
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, cur_file = None, "?"
    agg = {}
    for r in rows:
        if not r:
            continue
        if r[0] in ("File Path", "File Name"):
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) != len(hdr) or not r[0].isdigit():
            continue          # SASS rows (empty line number) are already summed into their source-line row
        ie, sm = num(r[hdr.index("Instructions Executed")]), num(r[hdr.index("# Samples")])
        a = agg.setdefault((cur_file, int(r[0])), [0.0, 0.0, r[1]])
        a[0] += ie
        a[1] += sm
    ti = sum(a[0] for a in agg.values()) or 1.0
    ts = sum(a[1] for a in agg.values()) or 1.0
    print(f"warp-inst {ti / 1e6:.1f}M  samples {ts:.0f}")
    for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{f:22s}:{ln:4d} inst {a[0] / 1e6:7.2f}M ({100 * a[0] / ti:4.1f}%) samp {100 * a[1] / ts:4.1f}%  {a[2].strip()[:80]}")


if __name__ == "__main__":
    main()
