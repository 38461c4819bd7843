"""SURVEY 8f rows measured: device-side tape validation (K5) and column materialisation (K6) on the
BASELINE config 2 / 3 inputs.  Prints one JSON line per kernel (GB/s of algorithmic bytes vs the measured
HBM peak, CPU baseline = the oracle's scalar statement, single thread).

    python tools/bench_tape.py [--size BYTES] [--steps K]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import csv_simd_b200 as cs  # noqa: E402
from oracle import oracle as O  # noqa: E402
from tools import gen  # noqa: E402


def peak():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "measured"
    except Exception:
        return 6650.0, "fallback"


def timed(stream, fn, steps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=1 << 30)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--cpu-rows", type=int, default=2_000_000)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    pk, pk_src = peak()
    ctx = cs.Context(0)
    stream = torch.cuda.current_stream(dev)
    ctx.set_stream(stream.cuda_stream)
    for wl, fcnt, crlf, fld in (("cfg2_unquoted", 16, False, 7), ("cfg3_quoted", 16, True, 1)):
        data, rows = gen.unquoted(a.size, seed=42) if wl == "cfg2_unquoted" else gen.quoted(a.size, seed=43)
        n = data.size
        d = torch.empty(n + 64, dtype=torch.uint8, device=dev)
        d[:n].copy_(torch.from_numpy(data))
        idx = ctx.index_build_device(d.data_ptr(), n)
        E = len(idx)
        rc, jump = idx.tape_init(fcnt, crlf)
        # ---- K5 ----
        rep = idx.tape_validate(fcnt, crlf)
        ms = timed(stream, lambda: idx.tape_validate(fcnt, crlf), a.steps)   # includes the 8-byte result read-back
        host = idx.to_host()
        t = time.perf_counter()
        want = O.tape_first_bad_slot(data, host, fcnt, crlf)
        cpu_s = time.perf_counter() - t
        assert rep["first_bad_slot"] == want
        alg = 8 * E + n
        print(json.dumps({
            "kernel": "tape_validate_kernel", "workload": wl, "metric": "csv_bytes_validated_per_sec",
            "value": n / (ms * 1e-3) / 1e9, "unit": "GB/s", "ms": ms, "index_entries": E, "ok": rep["ok"],
            "first_bad_slot": rep["first_bad_slot"],
            "roofline": {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": pk, "unit": "GB/s",
                         "frac": alg / (ms * 1e-3) / 1e9 / pk, "peak_source": pk_src,
                         "algorithmic_bytes": alg, "note": "8 B per index entry + every sector of the input"},
            "cpu_baseline": {"value": n / cpu_s / 1e9, "unit": "GB/s", "cores": 1, "kind": "port",
                             "sample": "whole input, scalar definition in oracle/csv_oracle.c"},
            "parity": "first_bad_slot equals the oracle's"}), flush=True)
        # ---- K6: one column of every record ----
        nrec = rc - 1
        flags = cs.FIELD_UNQUOTE | cs.FIELD_TRIM
        d_off = torch.empty(nrec + 1, dtype=torch.int64, device=dev)
        idx.materialize_column_device(fld, 0, nrec, flags, d_off.data_ptr(), 0, 0)
        torch.cuda.synchronize()
        total = int(d_off[-1].item())
        d_out = torch.empty(max(total, 1), dtype=torch.uint8, device=dev)
        ms = timed(stream, lambda: idx.materialize_column_device(fld, 0, nrec, flags, d_off.data_ptr(),
                                                                 d_out.data_ptr(), total), a.steps)
        # parity on a prefix the scalar oracle finishes in seconds + total length through the offsets
        k = min(nrec, a.cpu_rows)
        t = time.perf_counter()
        w_offs, w_out = O.materialize_column(data, host, rc, fcnt, crlf, fld, 0, k, flags)
        cpu_s = (time.perf_counter() - t) / 2      # the wrapper runs a sizing pass and a filling pass
        g_offs = d_off[:k + 1].cpu().numpy().view(np.uint64)
        assert (g_offs == w_offs).all() and d_out[:len(w_out)].cpu().numpy().tobytes() == w_out
        # raw bytes of the column = value bytes + quotes/escapes/padding removed; read side counted as the raw slices
        raw_bytes = int(((host[jump + fld + 1::jump][:nrec]).astype(np.int64) - host[jump + fld::jump][:nrec].astype(np.int64) - 1).sum())
        alg = nrec * (16 + 8) + raw_bytes + total
        print(json.dumps({
            "kernel": "materialize_offsets_kernel + materialize_write_kernel", "workload": wl, "field": fld,
            "metric": "values_materialised_per_sec", "value": nrec / (ms * 1e-3) / 1e6, "unit": "Mvalues/s", "ms": ms,
            "records": nrec, "value_bytes": total, "raw_bytes": raw_bytes, "flags": "UNQUOTE|TRIM",
            "roofline": {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": pk, "unit": "GB/s",
                         "frac": alg / (ms * 1e-3) / 1e9 / pk, "peak_source": pk_src, "algorithmic_bytes": alg,
                         "note": "per record 16 B of index + 8 B offset, raw slice read twice is NOT counted "
                                 "(once), value written once; strided gathers are sector-bound"},
            "cpu_baseline": {"value": k / cpu_s / 1e6, "unit": "Mvalues/s", "cores": 1, "kind": "port",
                             "sample": f"first {k} records, scalar definition in oracle/csv_oracle.c"},
            "parity": f"offsets and bytes of the first {k} records equal the oracle's"}), flush=True)
        # ---- K6 multi: 1, 4, 8 and all 16 columns in one sweep per pass (csvb200_materialize_columns_device) ----
        one_col_ms = ms
        for cols in ([fld], [1, 5, 9, 13], list(range(1, 16, 2)), list(range(16))):
            k_ = len(cols)
            d_offs = [torch.empty(nrec + 1, dtype=torch.int64, device=dev) for _ in cols]
            idx.materialize_columns_device(cols, 0, nrec, flags, [t_.data_ptr() for t_ in d_offs])
            torch.cuda.synchronize()
            totals = [int(t_[-1].item()) for t_ in d_offs]
            d_outs = [torch.empty(max(tt, 1), dtype=torch.uint8, device=dev) for tt in totals]
            ms_m = timed(stream, lambda: idx.materialize_columns_device(cols, 0, nrec, flags, [t_.data_ptr() for t_ in d_offs],
                                                                         [t_.data_ptr() for t_ in d_outs], totals), a.steps)
            # parity: every column equals the single-column kernel's output (which is pinned to the oracle above)
            ok = True
            for c_, f_ in enumerate(cols[:3]):
                o1 = torch.empty(nrec + 1, dtype=torch.int64, device=dev)
                idx.materialize_column_device(f_, 0, nrec, flags, o1.data_ptr(), 0, 0)
                v1 = torch.empty(max(totals[c_], 1), dtype=torch.uint8, device=dev)
                idx.materialize_column_device(f_, 0, nrec, flags, o1.data_ptr(), v1.data_ptr(), totals[c_])
                torch.cuda.synchronize()
                ok = ok and bool(torch.equal(o1, d_offs[c_])) and bool(torch.equal(v1[:totals[c_]], d_outs[c_][:totals[c_]]))
            assert ok
            print(json.dumps({
                "kernel": "materialize_multi_offsets_kernel + materialize_multi_write_kernel", "workload": wl, "columns": k_,
                "metric": "values_materialised_per_sec", "value": k_ * nrec / (ms_m * 1e-3) / 1e6, "unit": "Mvalues/s",
                "ms": ms_m, "ms_single_column_kernel": one_col_ms, "x_single_column": ms_m / one_col_ms,
                "ms_if_looped_over_single_column": k_ * one_col_ms, "value_bytes": sum(totals),
                "parity": "offsets and bytes of the first 3 columns equal the single-column kernels' output"}), flush=True)
            del d_offs, d_outs
        # ---- fused by-products: CSVB200_BUILD_VALIDATE against the plain build, same bytes ----
        def build_plain():
            ctx.index_build_device(d.data_ptr(), n).free()

        def build_val():
            ctx.index_build_device(d.data_ptr(), n, cs.BUILD_VALIDATE).free()
        ms_plain = timed(stream, build_plain, a.steps)
        ms_val = timed(stream, build_val, a.steps)
        iv = ctx.index_build_device(d.data_ptr(), n, cs.BUILD_VALIDATE)
        asc, newlines = iv.validation()
        t0_ = time.perf_counter()
        v_up = iv.validate_utf8()
        flagged_s = time.perf_counter() - t0_
        want_nl = int(np.isin(data[host[1:].astype(np.int64)], np.array([0x0D, 0x0A], dtype=np.uint8)).sum())
        assert asc == O.is_ascii(data) and newlines == want_nl and v_up is None
        iv.free()
        print(json.dumps({
            "kernel": "index_build_tma_kernel<validate>", "workload": wl, "metric": "csv_bytes_indexed_per_sec",
            "ms_build_plain": ms_plain, "ms_build_with_byproducts": ms_val, "overhead": ms_val / ms_plain - 1.0,
            "is_ascii": asc, "newlines_outside_quotes": newlines, "utf8_pass_over_flagged_tiles_ms": flagged_s * 1e3,
            "parity": "is_ascii == reader::is_ascii restatement; newline count == CR/LF entries of the oracle-checked index"}),
            flush=True)
        # ---- K7: ASCII / UTF-8 validation of the same bytes ----
        d_res = torch.empty(2, dtype=torch.int64, device=dev)
        ms = timed(stream, lambda: ctx.validate_utf8_device(d.data_ptr(), n, d_res.data_ptr()), a.steps)
        r = d_res.cpu().numpy().view(np.uint64)
        t = time.perf_counter()
        want_asc = O.is_ascii(data)
        cpu_s = time.perf_counter() - t
        assert int(r[0]) == 0xFFFFFFFFFFFFFFFF and bool(r[1] == 0) == want_asc
        print(json.dumps({
            "kernel": "utf8_validate_kernel", "workload": wl, "metric": "csv_bytes_validated_per_sec",
            "value": n / (ms * 1e-3) / 1e9, "unit": "GB/s", "ms": ms, "is_ascii": want_asc,
            "roofline": {"bound": "hbm", "achieved": n / (ms * 1e-3) / 1e9, "peak": pk, "unit": "GB/s",
                         "frac": n / (ms * 1e-3) / 1e9 / pk, "peak_source": pk_src, "algorithmic_bytes": int(n)},
            "cpu_baseline": {"value": n / cpu_s / 1e9, "unit": "GB/s", "cores": 1, "kind": "port",
                             "sample": "whole input, reader::is_ascii restatement (src/reader.rs:26-132)"},
            "parity": "is_ascii equals the restatement; valid_up_to = none"}), flush=True)
        idx.free()
        del d, d_off, d_out
        torch.cuda.empty_cache()
    # K7 on multi-byte text (the slow path of the kernel): every line carries 2-, 3- and 4-byte sequences
    line = "id,naïve café,日本語のテキスト,🙂🙃,1234\n".encode()
    reps = a.size // len(line)
    txt = np.frombuffer(line * reps, dtype=np.uint8)
    d = torch.from_numpy(txt.copy()).to(dev)
    d_res = torch.empty(2, dtype=torch.int64, device=dev)
    ms = timed(stream, lambda: ctx.validate_utf8_device(d.data_ptr(), txt.size, d_res.data_ptr()), a.steps)
    r = d_res.cpu().numpy().view(np.uint64)
    t = time.perf_counter()
    sample = txt[:64 << 20].tobytes()
    want = O.utf8_valid_up_to(sample[:len(sample) // len(line) * len(line)])
    cpu_s = time.perf_counter() - t
    assert int(r[0]) == 0xFFFFFFFFFFFFFFFF and want is None and int(r[1]) == 1
    print(json.dumps({
        "kernel": "utf8_validate_kernel", "workload": "multi-byte UTF-8 text (58 % non-ASCII bytes)",
        "metric": "csv_bytes_validated_per_sec", "value": txt.size / (ms * 1e-3) / 1e9, "unit": "GB/s", "ms": ms,
        "roofline": {"bound": "hbm", "achieved": txt.size / (ms * 1e-3) / 1e9, "peak": pk, "unit": "GB/s",
                     "frac": txt.size / (ms * 1e-3) / 1e9 / pk, "peak_source": pk_src, "algorithmic_bytes": int(txt.size)},
        "cpu_baseline": {"value": (64 << 20) / cpu_s / 1e9, "unit": "GB/s", "cores": 1, "kind": "port",
                         "sample": "first 64 MiB through CPython's strict UTF-8 decoder (the checker the tests use)"},
        "parity": "well-formed on both"}), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
