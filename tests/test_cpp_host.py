"""The C++ host mirror of the crate's API (csrc/host/csv_simd.hpp) driven through its flat C shim:
CPU tests for the host-only pieces, GPU tests for create / seek / chunks through libcsvb200."""
import ctypes as C
import os
import tempfile

import numpy as np
import pytest

from csv_simd_b200 import build as cbuild
from oracle import oracle as O
from tests import cases
from tests.conftest import golden_bytes

OK, IO, MISSING, INVALID_STATE, INVALID_CSV, PANIC, GPU = 0, 1, 2, 3, 4, 5, 6


@pytest.fixture(scope="module")
def shim():
    so = cbuild.build_host()
    L = C.CDLL(so)
    L.csvsimd_last_error.restype = C.c_char_p
    L.csvsimd_create.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
    L.csvsimd_tape_free.argtypes = [C.c_void_p]
    for f in ("record_cnt", "field_cnt", "record_offset"):
        getattr(L, "csvsimd_tape_" + f).argtypes = [C.c_void_p]
        getattr(L, "csvsimd_tape_" + f).restype = C.c_uint32
    L.csvsimd_tape_jump.argtypes = [C.c_void_p]
    L.csvsimd_tape_jump.restype = C.c_uint64
    L.csvsimd_tape_is_crlf.argtypes = [C.c_void_p]
    L.csvsimd_tape_index_len.argtypes = [C.c_void_p]
    L.csvsimd_tape_index_len.restype = C.c_size_t
    L.csvsimd_tape_index_copy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    L.csvsimd_tape_header_name.argtypes = [C.c_void_p, C.c_uint32, C.c_char_p, C.c_size_t]
    u64p, ip = C.POINTER(C.c_uint64), C.POINTER(C.c_int)
    L.csvsimd_tape_seek_record.argtypes = [C.c_void_p, C.c_uint32, u64p, u64p, ip]
    L.csvsimd_tape_seek_field.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, u64p, u64p, ip]
    L.csvsimd_tape_seek_fields.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    L.csvsimd_tape_chunks.argtypes = [C.c_void_p, C.c_uint8, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    L.csvsimd_boundaries.argtypes = [C.c_uint32, C.c_uint8, C.c_void_p, C.c_size_t]
    L.csvsimd_header.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), ip]
    L.csvsimd_core_seek_before_init.argtypes = [C.c_char_p]
    from csv_simd_b200._lib import TapeReport
    L.csvsimd_tape_validate.argtypes = [C.c_void_p, C.POINTER(TapeReport)]
    L.csvsimd_tape_column.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p,
                                      C.c_size_t, C.POINTER(C.c_size_t)]
    L.csvsimd_tape_utf8.argtypes = [C.c_void_p, u64p, ip]
    return L


def _tmp(data: bytes) -> str:
    f = tempfile.NamedTemporaryFile(suffix=".csv", delete=False)
    f.write(data)
    f.close()
    return f.name


# ---- CPU ------------------------------------------------------------------------------------------
def test_cpp_boundaries_vs_oracle(shim):
    out = np.zeros(2 * 256, dtype=np.uint64)
    for task, jobs in ((8, 3), (1000, 12), (8, 12), (0, 3), (7, 0), (65537, 255), (2 ** 32 - 1, 7)):
        n = shim.csvsimd_boundaries(task, jobs, out.ctypes.data, 256)
        want = O.boundaries(task, jobs)
        if want is None:
            assert n == 0
        else:
            assert [(int(out[2 * i]), int(out[2 * i + 1])) for i in range(n)] == want
    # src/tape.rs:362-384 doctest
    assert shim.csvsimd_boundaries(1000, 12, out.ctypes.data, 256) == 12 and (out[22], out[23]) == (917, 83)


def test_cpp_header_vs_oracle(shim):
    fc, ro, crlf = C.c_uint32(), C.c_uint32(), C.c_int()
    for name in ("sample.csv", "sample_rx.csv", "reader_test01.csv"):
        data = golden_bytes(name)
        buf = np.frombuffer(data, dtype=np.uint8)
        assert shim.csvsimd_header(buf.ctypes.data, buf.size, C.byref(fc), C.byref(ro), C.byref(crlf)) == OK
        want = O.header_new(data)
        assert (fc.value, ro.value, bool(crlf.value)) == (want.field_cnt, want.record_offset, want.crlf)
    for seed in range(100):
        data = cases.rand_bytes(30 + seed, 7000 + seed, weights=[5, 5, 5, 2, 1, 0.3, 0.3, 2, 1, 0.2, 0.2])
        buf = np.frombuffer(data, dtype=np.uint8)
        rc = shim.csvsimd_header(buf.ctypes.data, buf.size, C.byref(fc), C.byref(ro), C.byref(crlf))
        try:
            want = O.header_new(data)
        except O.OraclePanic:
            assert rc == PANIC
            continue
        assert rc == OK and (fc.value, ro.value, bool(crlf.value)) == (want.field_cnt, want.record_offset, want.crlf)


def test_cpp_create_missing_file_is_io(shim):
    t = C.c_void_p()
    assert shim.csvsimd_create(b"/nonexistent/missing.csv", C.byref(t)) == IO


# ---- GPU ------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["sample.csv", "sample_rx.csv"])
def test_cpp_create_and_seek(shim, golden, name):
    g = golden[name]
    data = golden_bytes(name)
    path = _tmp(data)
    try:
        t = C.c_void_p()
        assert shim.csvsimd_create(path.encode(), C.byref(t)) == OK, shim.csvsimd_last_error()
        assert shim.csvsimd_tape_record_cnt(t) == g["record_cnt"] and shim.csvsimd_tape_jump(t) == g["jump"]
        assert shim.csvsimd_tape_field_cnt(t) == g["field_cnt"] and bool(shim.csvsimd_tape_is_crlf(t)) == g["crlf"]
        assert shim.csvsimd_tape_record_offset(t) == g["record_offset"]
        n = shim.csvsimd_tape_index_len(t)
        idx = np.zeros(n, dtype=np.uint64)
        assert shim.csvsimd_tape_index_copy(t, idx.ctypes.data, n) == OK and idx.tolist() == g["index"]
        buf = C.create_string_buffer(256)
        names = []
        for i in range(g["field_cnt"]):
            assert shim.csvsimd_tape_header_name(t, i, buf, 256) == OK
            names.append(buf.value.decode())
        assert names == g["header"]
        s, ln, f = C.c_uint64(), C.c_uint64(), C.c_int()
        for r, want in g["seek_record"].items():
            assert shim.csvsimd_tape_seek_record(t, int(r), C.byref(s), C.byref(ln), C.byref(f)) == OK
            assert (data[s.value:s.value + ln.value].decode() if f.value else None) == want
        for key, want in g["seek_field"].items():
            r, fl = map(int, key.split(","))
            assert shim.csvsimd_tape_seek_field(t, r, fl, C.byref(s), C.byref(ln), C.byref(f)) == OK
            assert (data[s.value:s.value + ln.value].decode() if f.value else None) == want
        # batched seek through the K4 kernel == scalar host seeks
        keys = [tuple(map(int, k.split(","))) for k in g["seek_field"]]
        rec = np.array([k[0] for k in keys], dtype=np.uint32)
        fld = np.array([k[1] for k in keys], dtype=np.uint32)
        out = np.zeros((len(keys), 2), dtype=np.uint64)
        assert shim.csvsimd_tape_seek_fields(t, rec.ctypes.data, fld.ctypes.data, len(keys), out.ctypes.data) == OK
        for (a, b), want in zip(out.tolist(), g["seek_field"].values()):
            assert (None if a == 0xFFFFFFFFFFFFFFFF else data[a:b].decode()) == want
        # chunks vs the oracle restatement of Tape::chunks
        ch = np.zeros(4 * 16, dtype=np.uint64)
        nch = C.c_size_t()
        assert shim.csvsimd_tape_chunks(t, 4, ch.ctypes.data, 16, C.byref(nch)) == OK
        want = O.chunks(g["record_cnt"], g["jump"], 4)
        got = [dict(id=int(ch[4 * i]), start=int(ch[4 * i + 1]), end=int(ch[4 * i + 2]), record_cnt=int(ch[4 * i + 3]))
               for i in range(nch.value)]
        assert got == want
        shim.csvsimd_tape_free(t)
    finally:
        os.unlink(path)


@pytest.mark.gpu
def test_cpp_error_behaviour(shim):
    t = C.c_void_p()
    p = _tmp(golden_bytes("reader_test01.csv"))      # ragged: (E-1) % jump != 0
    try:
        assert shim.csvsimd_create(p.encode(), C.byref(t)) == INVALID_CSV
        assert b"Unsupported csv structure" in shim.csvsimd_last_error()
    finally:
        os.unlink(p)
    p = _tmp(b"a,b\n1,2\n")                           # n < 64: the reference panics in reader::read
    try:
        assert shim.csvsimd_create(p.encode(), C.byref(t)) == PANIC
    finally:
        os.unlink(p)
    p = _tmp(golden_bytes("sample.csv"))
    try:
        assert shim.csvsimd_core_seek_before_init(p.encode()) == INVALID_STATE   # record_source.rs:77-79
    finally:
        os.unlink(p)


@pytest.mark.gpu
def test_cpp_validate_column_utf8(shim):
    """The SURVEY 8f additions through the C++ mirror: Tape::validate, Tape::column, Tape::utf8_valid_up_to."""
    from csv_simd_b200._lib import TapeReport
    data = golden_bytes("sample_rx.csv")
    path = _tmp(data)
    try:
        t = C.c_void_p()
        assert shim.csvsimd_create(path.encode(), C.byref(t)) == OK, shim.csvsimd_last_error()
        rep = TapeReport()
        assert shim.csvsimd_tape_validate(t, C.byref(rep)) == OK
        assert rep.ok == 1 and rep.record_cnt == 8 and rep.jump == 9 and rep.first_bad_slot == 0xFFFFFFFFFFFFFFFF
        host = O.read_sse(data)
        for fld in (0, 2, 5, 7):
            w_offs, w_out = O.materialize_column(data, host, 8, 8, True, fld, 0, 7, 3)
            offs = np.zeros(8, dtype=np.uint64)
            buf = np.zeros(4096, dtype=np.uint8)
            total = C.c_size_t()
            assert shim.csvsimd_tape_column(t, fld, 0, 7, 3, offs.ctypes.data, buf.ctypes.data, buf.size, C.byref(total)) == OK
            assert (offs == w_offs).all() and buf[:total.value].tobytes() == w_out
        v, ok = C.c_uint64(), C.c_int()
        assert shim.csvsimd_tape_utf8(t, C.byref(v), C.byref(ok)) == OK
        want = O.utf8_valid_up_to(data)
        assert bool(ok.value) == (want is None) and (want is None or v.value == want)
        shim.csvsimd_tape_free(t)
    finally:
        os.unlink(path)
