"""CPU tests of the oracle itself: pins it against everything the reference's own tests hold
for this path, against the committed golden vectors, and against an independent closed form."""
import numpy as np
import pytest

from oracle import oracle as O
from tests import cases
from tests.conftest import golden_bytes


def test_reference_mk_index_pin():
    # src/reader.rs:318-327: index[1] == 4, index[last] == 95 on res/reader_test01.csv
    idx = O.read_sse(golden_bytes("reader_test01.csv"))
    assert idx[1] == 4
    assert idx[-1] == 95


def test_reference_blsr_identity():
    # src/lib.rs:139-152: 0b01011100 & (0b01011100 - 1) == 0b01011000
    assert O.blsr(0b01011100) == 0b01011000
    assert O.blsr(0) == 0


def test_reference_boundaries_doctest():
    # src/tape.rs:362-384
    r = O.boundaries(8, 3)
    assert r == [(0, 3), (3, 3), (6, 2)] and sum(l for _, l in r) == 8
    r = O.boundaries(1000, 12)
    assert r[0] == (0, 84) and r[1] == (84, 84) and r[11] == (917, 83)
    assert sum(l for _, l in r) == 1000
    assert O.boundaries(8, 12) == [(0, 8)]
    assert O.boundaries(0, 3) is None


@pytest.mark.parametrize("name", ["reader_test01.csv", "sample.csv", "sample_rx.csv"])
def test_golden_vectors(golden, name):
    data = golden_bytes(name)
    g = golden[name]
    assert len(data) == g["n"]
    idx = O.read_sse(data)
    assert idx.tolist() == g["index"]
    assert O.read_closed_form(data)[0].tolist() == g["index"]
    assert O.closed_form_numpy(data).tolist() == g["index"]
    h = O.header_new(data)
    assert (h.header, h.crlf, h.field_cnt, h.record_offset) == (g["header"], g["crlf"], g["field_cnt"],
                                                                g["record_offset"])
    if g["tape_ok"]:
        assert O.tape_init(len(idx), h.field_cnt, h.crlf) == (g["jump"], g["record_cnt"])
    else:
        with pytest.raises(O.InvalidCsvFormat):
            O.tape_init(len(idx), h.field_cnt, h.crlf)


@pytest.mark.parametrize("name", ["sample.csv", "sample_rx.csv"])
def test_golden_seeks(golden, name):
    data = golden_bytes(name)
    g = golden[name]
    idx = np.asarray(g["index"], dtype=np.uint64)
    for r, want in g["seek_record"].items():
        rg = O.seek_record(idx, len(data), g["record_cnt"], g["jump"], g["field_cnt"], int(r))
        got = None if rg is None else data[rg[0]:rg[1]].decode()
        assert got == want
    for key, want in g["seek_field"].items():
        r, f = map(int, key.split(","))
        rg = O.seek_field(idx, len(data), g["record_cnt"], g["field_cnt"], g["crlf"], r, f)
        got = None if rg is None else data[rg[0]:rg[1]].decode()
        assert got == want


def test_survey_known_answers(golden):
    # values derived independently during the survey
    g = golden["sample.csv"]
    assert len(g["index"]) == 46 and g["record_cnt"] == 15 and g["jump"] == 3
    assert g["seek_record"]["0"] == 'Edm nd,3, "o"' and g["seek_field"]["6,0"] == "iharlotte"
    g = golden["sample_rx.csv"]
    assert len(g["index"]) == 73 and g["record_cnt"] == 8 and g["jump"] == 9 and g["record_offset"] == 128
    assert g["seek_field"]["1,2"] == '"INTERNAL MED, CARD. ELECTROPHYSIOLOGY"'
    assert golden["reader_test01.csv"]["tape_ok"] is False and len(golden["reader_test01.csv"]["index"]) == 17


@pytest.mark.parametrize("name,data", cases.edge_cases(), ids=[c[0] for c in cases.edge_cases()])
def test_sse_restatement_equals_closed_form(name, data):
    a = O.read_sse(data)
    b, _ = O.read_closed_form(data)
    assert a.shape == b.shape and (a == b).all()
    assert (O.closed_form_numpy(data) == b).all()


def test_fuzz_sse_vs_closed_form():
    for seed in range(300):
        n = 64 + (seed * 37) % 700
        data = cases.rand_bytes(n, seed) if seed % 3 else cases.full_random(n, seed)
        a = O.read_sse(data)
        b, _ = O.read_closed_form(data)
        assert (a == b).all(), seed


def test_small_inputs_panic_in_reference():
    for _, data in cases.small_cases():
        with pytest.raises(O.OraclePanic):
            O.read_sse(data)


def test_class_bytes_all_256():
    # LUT enumeration (src/stage1.rs:24-35): only six byte values classify
    data = bytes(range(256))
    got = np.concatenate([O.structure_run(data, at) for at in range(0, 256, 16)])
    want = np.zeros(256, dtype=np.uint8)
    want[0x0A] = want[0x0D] = 1
    want[0x2C] = 2
    want[0x20] = 4
    want[0x5C] = 8
    want[0x22] = 16
    assert (got == want).all()


def test_shard_summary_composes():
    data = cases.rand_bytes(5000, 11)
    full, _ = O.read_closed_form(data)
    cuts = [0, 13, 1700, 1701, 3333, 5000]
    par, segs = 0, []
    for k in range(len(cuts) - 1):
        seg = data[cuts[k]:cuts[k + 1]]
        p, c0, s = O.shard_summary(seg)
        idx, endp = O.read_closed_form(seg, start_parity=par, pos_bias=cuts[k], with_sentinel=(k == 0))
        assert len(idx) - (1 if k == 0 else 0) == (s - c0 if par else c0)
        assert endp == par ^ p
        segs.append(idx)
        par ^= p
    assert (np.concatenate(segs) == full).all()


def test_chunks_restatement():
    ch = O.chunks(15, 3, 4)
    assert [c["record_cnt"] for c in ch] == [3, 4, 4, 3] and ch[0]["start"] == 3 and ch[-1]["end"] == 45
    assert O.chunks(0, 3, 4) is None


# ---- definitions beyond the reference: known answers of the scalar statements --------
def test_field_value_known_answers():
    U, T = 1, 2
    cases_ = [
        (b'plain', 0, b'plain'), (b'"quoted"', U, b'quoted'), (b'"quoted"', 0, b'"quoted"'),
        (b'  padded\t', T, b'padded'), (b'  padded\t', U, b'  padded\t'),
        (b'"with ""escapes"" inside"', U, b'with "escapes" inside'), (b'""', U, b''), (b'"', U, b'"'),
        (b'', U | T, b''), (b' "q, and\nnewline" ', U | T, b'q, and\nnewline'), (b' "x" ', U, b' "x" '),
        (b'""""', U, b'"'), (b'"a"""', U, b'a"'), (b'x"y', U, b'x"y'), (b'"a"b"', U, b'a"b'), (b'\t \t', T, b''),
        (b'" "', U | T, b' '),
    ]
    for raw, flags, want in cases_:
        assert O.field_value(raw, flags) == want, (raw, flags)


def test_tape_first_bad_slot_known_answers():
    d = b"a,b\n1,2\n3\n4,5\n"
    idx = O.closed_form_numpy(d)
    assert O.tape_first_bad_slot(d, idx, 2, False) == 5            # the '\n' after "3" sits in a ',' slot
    good = b"a,b\r\n1,2\r\n"
    assert O.tape_first_bad_slot(good, O.closed_form_numpy(good), 2, True) == 0xFFFFFFFFFFFFFFFF
    lone = b"a,b\r\n1,2\n\n"                                      # LF LF where CR LF is required
    assert O.tape_first_bad_slot(lone, O.closed_form_numpy(lone), 2, True) == 5
    for name, fc, crlf in (("sample.csv", 3, False), ("sample_rx.csv", 8, True)):
        from tests.conftest import golden_bytes
        raw = golden_bytes(name)
        assert O.tape_first_bad_slot(raw, O.read_sse(raw), fc, crlf) == 0xFFFFFFFFFFFFFFFF
    raw = golden_bytes("reader_test01.csv")                        # ragged last row
    assert O.tape_first_bad_slot(raw, O.read_sse(raw), 3, False) != 0xFFFFFFFFFFFFFFFF


def test_is_ascii_restatement():
    rng = np.random.default_rng(11)
    for n in list(range(0, 40)) + [63, 64, 65, 1000, 4097]:
        base = rng.integers(0, 128, size=n, dtype=np.uint8)
        for off in (0, 1, 3, 7):           # every alignment of the slice start
            buf = np.zeros(n + 16, dtype=np.uint8)
            buf[off:off + n] = base
            view = buf[off:off + n]
            assert O.is_ascii(view) is True
            for hit in {0, n // 2, n - 1} if n else ():
                view[hit] |= 0x80
                assert O.is_ascii(view) is False, (n, off, hit)
                view[hit] &= 0x7F


def test_utf8_valid_up_to_known_answers():
    assert O.utf8_valid_up_to(b"plain ascii") is None
    assert O.utf8_valid_up_to("héllo, wörld – ✓ 🙂".encode()) is None
    assert O.utf8_valid_up_to(b"ab\xe2\x82") == 2            # truncated at the end
    assert O.utf8_valid_up_to(b"ab\xc0\xaf") == 2            # overlong lead C0
    assert O.utf8_valid_up_to(b"\xed\xa0\x80") == 0          # surrogate
    assert O.utf8_valid_up_to(b"a\x80") == 1                 # stray continuation
    assert O.utf8_valid_up_to(b"\xf4\x90\x80\x80") == 0      # > U+10FFFF
