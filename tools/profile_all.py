"""One launch (a few for the hot kernel) of EVERY kernel in libcsvb200 on a mid-sized input, for ncu:
    python tools/profile_all.py > gpurun_out/plain.log && \
    ncu --set full --clock-control none --import-source on -o gpurun_out/prof_all python tools/profile_all.py
usage: python tools/profile_all.py [bytes]   (default 256 MiB per input)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import csv_simd_b200 as cs  # noqa: E402
from tools import gen  # noqa: E402


def main():
    size = int(sys.argv[1]) if len(sys.argv) > 1 else (256 << 20)
    dev = torch.device("cuda", 0)
    ctx = cs.Context(0)
    os.environ["CSVB200_KERNEL"] = "simple"
    ctx_simple = cs.Context(0)
    del os.environ["CSVB200_KERNEL"]
    for wl, fcnt, crlf in (("cfg2_unquoted", 16, False), ("cfg3_quoted", 16, True)):
        data, rows = gen.unquoted(size, seed=42) if wl == "cfg2_unquoted" else gen.quoted(size, seed=43)
        n = data.size
        d = torch.empty(n + 64, dtype=torch.uint8, device=dev)
        d[:n].copy_(torch.from_numpy(data))
        # K1-K3 fused: the TMA pipeline (hot kernel) and the one-tile-per-CTA kernel
        idx = ctx.index_build_device(d.data_ptr(), n)
        E = len(idx)
        i2 = ctx_simple.index_build_device(d.data_ptr(), n)
        assert len(i2) == E
        i2.free()
        rc, jump = idx.tape_init(fcnt, crlf)
        # multi-GPU pieces on one device: pass A, predictor, speculative build + verify (+ conditional rebuild)
        par = torch.zeros(1, dtype=torch.int32, device=dev)
        ctx.shard_quote_parity_device(d.data_ptr(), n, par.data_ptr())
        res = torch.zeros((2, 4), dtype=torch.int64, device=dev)
        half = (n // 2) | 5
        tail = d[half:half + (n - half)].clone()
        torch.cuda.synchronize()   # torch's stream and the context's own (non-blocking) stream are not ordered otherwise
        a = ctx.index_build_shard_speculative(d.data_ptr(), half, 0, 0, True, res[0].data_ptr())
        b = ctx.index_build_shard_speculative(tail.data_ptr(), n - half, 1, half, False, res[1].data_ptr())
        fin = torch.zeros((2, 2), dtype=torch.int64, device=dev)
        torch.cuda.synchronize()
        a.shard_verify(res.data_ptr(), 2, fin.data_ptr())
        b.shard_verify(res.data_ptr(), 2, fin.data_ptr())
        assert len(a) + len(b) == E, (len(a), len(b), E, res.tolist(), fin.tolist(), a.shard_redone(), b.shard_redone())
        a.free()
        b.free()
        # K4 lookups + byte gather
        rec, fld = gen.queries(2_000_000, rc, fcnt, seed=46)
        d_rec, d_fld = torch.from_numpy(rec.view(np.int32)).to(dev), torch.from_numpy(fld.view(np.int32)).to(dev)
        d_out = torch.empty((rec.size, 2), dtype=torch.int64, device=dev)
        idx.seek_fields_device(d_rec.data_ptr(), d_fld.data_ptr(), rec.size, d_out.data_ptr())
        idx.seek_records_device(d_rec.data_ptr(), rec.size, d_out.data_ptr())
        idx.gather_fields(rec[:200_000], fld[:200_000])
        # K5 / K6 / K7
        idx.tape_validate(fcnt, crlf)
        nrec = rc - 1
        d_off = torch.empty(nrec + 1, dtype=torch.int64, device=dev)
        idx.materialize_column_device(1, 0, nrec, 3, d_off.data_ptr(), 0, 0)
        torch.cuda.synchronize()
        total = int(d_off[-1].item())
        d_val = torch.empty(max(total, 1), dtype=torch.uint8, device=dev)
        idx.materialize_column_device(1, 0, nrec, 3, d_off.data_ptr(), d_val.data_ptr(), total)
        d_res = torch.empty(2, dtype=torch.int64, device=dev)
        ctx.validate_utf8_device(d.data_ptr(), n, d_res.data_ptr())
        torch.cuda.synchronize()
        # round 2: several columns in one sweep; the build with its by-products + the flagged-tile UTF-8 pass
        cols = [1, 5, 9, 13]
        d_offs = [torch.empty(nrec + 1, dtype=torch.int64, device=dev) for _ in cols]
        idx.materialize_columns_device(cols, 0, nrec, 3, [t.data_ptr() for t in d_offs])
        torch.cuda.synchronize()
        totals = [int(t[-1].item()) for t in d_offs]
        d_vals = [torch.empty(max(t, 1), dtype=torch.uint8, device=dev) for t in totals]
        idx.materialize_columns_device(cols, 0, nrec, 3, [t.data_ptr() for t in d_offs], [t.data_ptr() for t in d_vals], totals)
        dv = d.clone()
        dv[n // 3:n // 3 + 4] = torch.tensor([0xF0, 0x9F, 0x99, 0x82], dtype=torch.uint8, device=dev)   # one emoji: one flagged tile
        torch.cuda.synchronize()
        iv = ctx.index_build_device(dv.data_ptr(), n, cs.BUILD_VALIDATE)
        iv.validation()
        iv.validate_utf8()
        iv.free()
        # round 2: the exchange inside the launch (two ranks emulated in rank order on this GPU) and the lookups over
        # a distributed index (segments on "two devices" = this GPU twice)
        c2 = [cs.Context(0), cs.Context(0)]
        e2 = [c.exchange(k, 2) for k, c in enumerate(c2)]
        cs.Exchange.connect_local(e2)
        xa = c2[0].index_build_shard_exchange(e2[0], d.data_ptr(), half, 0)
        xa.sync()
        xb = c2[1].index_build_shard_exchange(e2[1], tail.data_ptr(), n - half, half)
        assert len(xa) + len(xb) == E
        xa.free()
        xb.free()
        for e in e2:
            e.close()
        for c in c2:
            c.close()
        idx.free()
        print(f"{wl}: n={n} E={E} ok")
    m = cs.Multi([0, 0])
    mi = m.index_build_distributed(data[:64 << 20])
    try:
        mi.tape_init(16, True)      # the prefix ends mid-row: InvalidCsvFormat is reported AFTER the metadata is set
    except cs.InvalidCsvFormat:
        pass
    rec, fld = gen.queries(1_000_000, 100000, 16, seed=46)
    mi.seek_fields(rec, fld)
    mi.free()
    m.close()
    # K1 known-answer exports on a small input
    small = data[:1 << 20]
    ctx.block_masks(small)
    ctx.class_bytes(small)
    ctx.close()
    ctx_simple.close()


if __name__ == "__main__":
    main()
