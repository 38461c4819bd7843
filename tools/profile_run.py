"""Small fixed workload for ncu: a few device-resident index builds of one BASELINE config.
usage: python tools/profile_run.py [cfg2_unquoted|cfg3_quoted] [builds] [bytes]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import csv_simd_b200 as cs  # noqa: E402
from tools import gen  # noqa: E402


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2_unquoted"
    builds = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    size = int(sys.argv[3]) if len(sys.argv) > 3 else (1 << 30)
    data, _ = (gen.unquoted(size, seed=42) if wl == "cfg2_unquoted" else gen.quoted(size, seed=43))
    dev = torch.device("cuda", 0)
    d = torch.empty(data.size + 64, dtype=torch.uint8, device=dev)
    d[:data.size].copy_(torch.from_numpy(data))
    ctx = cs.Context(0)
    ms = []
    for _ in range(builds):
        idx = ctx.index_build_device(d.data_ptr(), data.size)
        E = len(idx)
        ms.append(ctx.last_build_ms())
        idx.free()
    print(f"{wl}: n={data.size} E={E} kernel_ms={['%.3f' % m for m in ms]}")
    ctx.close()


if __name__ == "__main__":
    main()
