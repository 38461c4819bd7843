"""profiles/r02_kvariants.md from the kv*.jsonl lines of tools/kvariants.py (gpurun_out/, copied to profiles/ as
r02_kvariants_raw.jsonl).  The meaning of CSVB200_TUNE changed twice during the round, so the table is keyed by run."""
import glob
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHAPES = {0x1000: "A: 2 CTAs/SM, 95 regs, 2 staging buffers (round 1)", 0x2000: "B: 3 CTAs/SM, 64 regs, 1 staging buffer",
          0x3000: "D: A with 128 KiB per descriptor (kSub=4)", 0x4000: "E: B with kSub=4 (spills)",
          0x5000: "F: B with 2 super-tiles of skew (spills)", 0x6000: "G: A with 2 super-tiles of skew",
          0x7000: "H: 2 CTAs/SM x 12 worker warps (96 KiB per descriptor)", 0x8000: "I: 1 CTA/SM x 16 worker warps (128 KiB per descriptor)",
          0x9000: "J: B without skew", 0xC000: "M: B with the transpose's right shifts on the FMA pipe (__umulhi)", 0xD000: "N: B with the tail-less expansion loop on every sub-tile", 0xE000: "O: B with the tail-less expansion loop on sparse sub-tiles (kv15; the default from kv16 on, where 0xE000 = P: B with the round-1 expansion loop)", 0xA000 if False else 0x10000: "", 0xF000: "Q: default + sparse expansion loop from the highest bit down (kv16) / first two entries of a group without a loop (kv17)", 0xA000: "K: B with 32 KiB per descriptor (kSub=1)", 0xB000: "L: K without skew", 0: "default (B)"}


def label(run: int, t: int) -> str:
    if run == 2:   # first encoding: bits 8-9
        return SHAPES[0x1000] if t == 256 else SHAPES[0x2000]
    shape, m = t & 0xF000, t & 0xFFF
    nm = SHAPES.get(shape, hex(shape))
    if run == 18 and shape in (0xA000, 0xB000):
        nm = "B with %d descriptors per look-back round trip" % (128 if shape == 0xA000 else 64)
    if run == 22 and shape == 0xA000:
        nm = "B + forward push (a resolved tile publishes the prefixes of the tiles after it whose aggregates are in)"
    if run == 19 and shape == 0xA000:
        nm = "B with the two-level look-back (a tile without a prefix in its window publishes the composite of 33 tiles)"
    if run <= 7:   # before the just-in-time tickets: low bits = look-back knobs
        if m in (1, 2, 3):
            nm += " + " + {1: "re-poll one descriptor", 2: "look-back window 64", 3: "look-back window 128"}[m]
    else:          # kv8+: tickets are just-in-time by default; bit 4 = round-1 ticket policy, bits 8 / 16 = where `go` is signalled
        nm += " + JIT tickets" if not (m & 4) else " + tickets drawn when a ring slot frees (round 1)"
        if 18 <= run <= 20 and (m & 3):
            nm += {1: ", re-poll one descriptor (the default after kv20)", 2: ", every lane waits for its own descriptor", 3: "?"}[m & 3]
        if run >= 21:
            nm += ", re-poll the whole window (the default up to kv20)" if (m & 1) else ", re-poll one descriptor"
        if (m & 24) == 8:
            nm += ", go right after the prefix wait"
        if (m & 24) == 16:
            nm += ", go after the whole compaction"
        if run == 12 and (m & 32):
            nm += ", copy-out with one 32-bit add per entry"
        if run == 13 and (m & 0xE0):
            nm += f", service-warp waits back off {32 << ((m >> 5) & 3)} ns after a failed try ({'look-back and producer' if m & 128 else 'look-back warp'})"
    if m & 0x400:
        nm += " + NO CHAIN (timing experiment: every look-back answered on its first poll)"
    return nm


def main():
    out = ["# Round 2: shapes and knobs of `index_build_tma_kernel`, A/B on one box per run", "",
           "`tools/kvariants.py`: all variants of a run in ONE process on the same bytes, interleaved round-robin, 15-20 timed builds each,",
           "index compared element-wise with the first variant's (which the `-m gpu` suite pins to the oracle).  Times are CUDA events around",
           "the launch; frac = (N + 8E) / t / 6453.4 GB/s (measured copy peak).  Runs are on different boxes: compare within a run.", "",
           "| run | workload | variant | kernel ms (avg / min) | frac of measured HBM | CSV GB/s | index == first variant |", "|---|---|---|---|---|---|---|"]
    raw = []
    files = sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "kv*.jsonl")), key=lambda f: int(os.path.basename(f)[2:-6]))
    for f in files:
        run = int(os.path.basename(f)[2:-6])
        if run == 1:
            continue
        for line in open(f):
            d = json.loads(line)
            if run == 3 and (d["tune"] & 0x400):
                continue   # the invalid first attempt at a no-chain bound (fake prefixes changed the output)
            d["run"] = run
            raw.append(d)
            out.append(f"| kv{run} | {d['workload'][:4]} | {label(run, d['tune'])} | {d['kernel_ms_avg']:.4f} / {d['kernel_ms_min']:.4f} | "
                       f"{d['frac_of_measured_hbm']:.3f} | {d['csv_gbs']:.0f} | {d['parity_vs_first']} |")
    out += ["", "kv1 (not listed): shape C, a ring of three 16 KiB half-stages shared by two warp groups, gave an entry count off by 2 on the",
            "1 GiB cfg2 input (the `-m gpu` suite under that shape: 1 failed) and was dropped -- see DESIGN.md 4.1."]
    open(os.path.join(ROOT, "profiles", "r02_kvariants.md"), "w").write("\n".join(out) + "\n")
    with open(os.path.join(ROOT, "profiles", "r02_kvariants_raw.jsonl"), "w") as fo:
        for d in raw:
            fo.write(json.dumps(d) + "\n")
    print(len(raw), "rows")


if __name__ == "__main__":
    main()
