"""Regenerates tests/golden/golden.json from the CPU oracle.

The three CSV fixtures are byte-for-byte copies of the reference's res/ files
(res/sample.csv, res/sample_rx.csv, res/reader_test01.csv).  The reference cannot be run here
(Rust, no toolchain), so the vectors are produced by oracle/csv_oracle.c -- the literal SSE
restatement -- cross-checked against the independent closed form, and against the only values
the reference's own tests pin (src/reader.rs:318-327: index[1] == 4, index[last] == 95 on
reader_test01.csv).  Run:  python tests/golden/make_golden.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as O  # noqa: E402

SEEKS = {
    "sample.csv": dict(records=[0, 13, 14], fields=[(0, 0), (0, 2), (6, 0), (6, 1), (6, 2), (6, 3), (10, 1)]),
    "sample_rx.csv": dict(records=[0, 6, 7, 10], fields=[(1, 2), (6, 2), (6, 3), (6, 7), (6, 8), (10, 1), (6, 0), (6, 1)]),
}


def main():
    out = {}
    for name in ("reader_test01.csv", "sample.csv", "sample_rx.csv"):
        data = open(os.path.join(HERE, name), "rb").read()
        idx = O.read_sse(data)
        cf, _ = O.read_closed_form(data)
        assert (idx == cf).all() and (idx == O.closed_form_numpy(data)).all()
        h = O.header_new(data)
        ent = dict(n=len(data), index=[int(v) for v in idx], header=h.header, crlf=h.crlf,
                   field_cnt=h.field_cnt, record_offset=h.record_offset)
        try:
            jump, rc = O.tape_init(len(idx), h.field_cnt, h.crlf)
            ent.update(jump=jump, record_cnt=rc, tape_ok=True)
        except O.InvalidCsvFormat:
            ent.update(tape_ok=False)
        if name in SEEKS and ent["tape_ok"]:
            recs, flds = {}, {}
            for r in SEEKS[name]["records"]:
                rg = O.seek_record(idx, len(data), ent["record_cnt"], ent["jump"], h.field_cnt, r)
                recs[str(r)] = None if rg is None else data[rg[0]:rg[1]].decode()
            for r, f in SEEKS[name]["fields"]:
                rg = O.seek_field(idx, len(data), ent["record_cnt"], h.field_cnt, h.crlf, r, f)
                flds[f"{r},{f}"] = None if rg is None else data[rg[0]:rg[1]].decode()
            ent.update(seek_record=recs, seek_field=flds)
        out[name] = ent
    # reference-pinned values (src/reader.rs:325-326)
    assert out["reader_test01.csv"]["index"][1] == 4 and out["reader_test01.csv"]["index"][-1] == 95
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote golden.json")


if __name__ == "__main__":
    main()
