"""ctypes front-end of the CPU oracle -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product package (csv_simd_b200) never does.

Two independent models live here:
  * the C restatement in csv_oracle.c (literal SSE sequence of the reference,
    src/reader.rs:150-306 -> src/avx/stage1.rs:193-407 -> src/stage1.rs:162-296),
  * `closed_form_numpy`, a vectorised numpy statement of the closed form
    index = [0] ++ [i : b[i] in {',',CR,LF} and #quotes before i is even].
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libcsv_oracle.so")

OK, ERR_PANIC, ERR_UNALIGNED, ERR_OOM, ERR_INVALID_CSV_FORMAT = 0, 1, 2, 3, 4


class OraclePanic(RuntimeError):
    """The reference would panic (or hit UB) on this input."""


class InvalidCsvFormat(RuntimeError):
    """StructureError::InvalidCsvFormat (src/error.rs:19-20)."""


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "csv_oracle.c")
    hdr = os.path.join(_HERE, "csv_oracle.h")
    stale = (not os.path.exists(_SO)) or any(
        os.path.getmtime(p) > os.path.getmtime(_SO) for p in (src, hdr))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "libcsv_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return _SO


class _Header(C.Structure):
    _fields_ = [("field_cnt", C.c_uint32), ("record_offset", C.c_uint32),
                ("crlf", C.c_int), ("header_start", C.c_size_t),
                ("header_end", C.c_size_t)]


class _Boundary(C.Structure):
    _fields_ = [("start", C.c_uint64), ("len", C.c_uint64)]


class _Chunk(C.Structure):
    _fields_ = [("id", C.c_uint8), ("start", C.c_uint64), ("end", C.c_uint64),
                ("record_cnt", C.c_uint32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        u8p, u64p, szp = C.POINTER(C.c_uint8), C.POINTER(C.c_uint64), C.POINTER(C.c_size_t)
        L.oracle_read_sse.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(u64p), szp]
        L.oracle_read_sse_timed.argtypes = [C.c_void_p, C.c_size_t, szp, u64p]
        L.oracle_read_closed_form.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_uint64,
                                              C.c_int, C.POINTER(u64p), szp, C.POINTER(C.c_int)]
        L.oracle_shard_summary.argtypes = [C.c_void_p, C.c_size_t, u64p, u64p, u64p]
        L.oracle_shard_summary.restype = None
        L.oracle_free.argtypes = [C.c_void_p]
        L.oracle_free.restype = None
        L.oracle_structure_run.argtypes = [C.c_void_p, C.c_size_t, u8p]
        L.oracle_structure_run.restype = None
        L.oracle_header_new.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(_Header)]
        L.oracle_header_name.argtypes = [C.c_void_p, C.POINTER(_Header), C.c_uint32, szp, szp]
        L.oracle_tape_init.argtypes = [C.c_size_t, C.c_uint32, C.c_int, u64p, C.POINTER(C.c_uint32)]
        L.oracle_seek_record.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_uint32, C.c_uint64,
                                         C.c_uint32, C.c_uint32, u64p, u64p, C.POINTER(C.c_int)]
        L.oracle_seek_field.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_uint32, C.c_uint32,
                                        C.c_int, C.c_uint32, C.c_uint32, u64p, u64p, C.POINTER(C.c_int)]
        L.oracle_seek_fields_timed.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_uint32, C.c_uint32, C.c_int,
                                               C.c_void_p, C.c_void_p, C.c_size_t, u64p, u64p]
        L.oracle_boundaries.argtypes = [C.c_uint32, C.c_uint8, C.POINTER(_Boundary)]
        L.oracle_chunks.argtypes = [C.c_uint32, C.c_uint64, C.c_uint8, C.POINTER(_Chunk)]
        L.oracle_tape_first_bad_slot.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_uint32, C.c_int]
        L.oracle_tape_first_bad_slot.restype = C.c_uint64
        L.oracle_field_value.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_void_p]
        L.oracle_field_value.restype = C.c_size_t
        L.oracle_materialize_column.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32,
                                                C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p,
                                                C.c_void_p]
        L.oracle_materialize_column.restype = C.c_uint64
        L.oracle_is_ascii.argtypes = [C.c_void_p, C.c_size_t]
        L.oracle_blsr.argtypes = [C.c_uint64]
        L.oracle_blsr.restype = C.c_uint64
        _lib = L
    return _lib


def _check(rc: int):
    if rc == OK:
        return
    if rc == ERR_PANIC:
        raise OraclePanic("reference would panic here")
    if rc == ERR_INVALID_CSV_FORMAT:
        raise InvalidCsvFormat("Unsupported csv structure: likely variable number of fields")
    raise RuntimeError(f"oracle error {rc}")


def aligned_copy(data, align: int = 64) -> np.ndarray:
    """Copy `data` into a fresh numpy buffer whose base address is `align`-aligned
    (stands in for the page-aligned mmap the reference reads from)."""
    a = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
    raw = np.empty(a.size + align, dtype=np.uint8)
    off = (-raw.ctypes.data) % align
    out = raw[off:off + a.size]
    out[:] = a
    return out


def _take(ptr, n) -> np.ndarray:
    out = np.ctypeslib.as_array(ptr, shape=(n,)).copy() if n else np.zeros(0, dtype=np.uint64)
    lib().oracle_free(ptr)
    return out


def read_sse(data) -> np.ndarray:
    """reader::read, literal SSE restatement (needs n >= 64 like the reference)."""
    buf = aligned_copy(data)
    p = C.POINTER(C.c_uint64)()
    n = C.c_size_t()
    _check(lib().oracle_read_sse(buf.ctypes.data, buf.size, C.byref(p), C.byref(n)))
    return _take(p, n.value)


def read_sse_timed(buf: np.ndarray):
    """Timed leg for the CPU baseline: buf must already be 16-byte aligned."""
    n = C.c_size_t()
    s = C.c_uint64()
    _check(lib().oracle_read_sse_timed(buf.ctypes.data, buf.size, C.byref(n), C.byref(s)))
    return n.value, s.value


def read_closed_form(data, start_parity: int = 0, pos_bias: int = 0, with_sentinel: bool = True):
    a = np.ascontiguousarray(np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data)
    p = C.POINTER(C.c_uint64)()
    n = C.c_size_t()
    ep = C.c_int()
    _check(lib().oracle_read_closed_form(a.ctypes.data, a.size, int(start_parity), int(pos_bias),
                                         int(with_sentinel), C.byref(p), C.byref(n), C.byref(ep)))
    return _take(p, n.value), ep.value


def closed_form_numpy(data, start_parity: int = 0, pos_bias: int = 0, with_sentinel: bool = True) -> np.ndarray:
    a = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
    isq = a == 0x22
    before = np.cumsum(isq, dtype=np.int64) - isq  # quotes strictly before i
    sep = (a == 0x2C) | (a == 0x0D) | (a == 0x0A)
    keep = sep & (((before + start_parity) & 1) == 0)
    pos = np.flatnonzero(keep).astype(np.uint64) + np.uint64(pos_bias)
    if with_sentinel:
        pos = np.concatenate([np.zeros(1, dtype=np.uint64), pos])
    return pos


def shard_summary(data):
    a = np.ascontiguousarray(np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data)
    p, c0, s = C.c_uint64(), C.c_uint64(), C.c_uint64()
    lib().oracle_shard_summary(a.ctypes.data, a.size, C.byref(p), C.byref(c0), C.byref(s))
    return p.value, c0.value, s.value


def structure_run(chunk: bytes, at: int = 0) -> np.ndarray:
    a = np.ascontiguousarray(np.frombuffer(chunk, dtype=np.uint8))
    assert a.size - at >= 16
    out = np.zeros(16, dtype=np.uint8)
    lib().oracle_structure_run(a.ctypes.data, at, out.ctypes.data_as(C.POINTER(C.c_uint8)))
    return out


@dataclass
class Header:
    header: list
    crlf: bool
    field_cnt: int
    record_offset: int


def header_new(data) -> Header:
    a = np.ascontiguousarray(np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data)
    h = _Header()
    _check(lib().oracle_header_new(a.ctypes.data, a.size, C.byref(h)))
    names = []
    raw = a.tobytes()
    for i in range(h.field_cnt):
        s, e = C.c_size_t(), C.c_size_t()
        _check(lib().oracle_header_name(a.ctypes.data, C.byref(h), i, C.byref(s), C.byref(e)))
        names.append(raw[s.value:e.value].decode("utf-8", "replace"))
    return Header(names, bool(h.crlf), h.field_cnt, h.record_offset)


def tape_init(index_len: int, field_cnt: int, crlf: bool):
    j, r = C.c_uint64(), C.c_uint32()
    _check(lib().oracle_tape_init(index_len, field_cnt, int(crlf), C.byref(j), C.byref(r)))
    return j.value, r.value


def seek_record(index: np.ndarray, data_len: int, record_cnt: int, jump: int, field_cnt: int, r: int):
    s, e, f = C.c_uint64(), C.c_uint64(), C.c_int()
    _check(lib().oracle_seek_record(index.ctypes.data, index.size, data_len, record_cnt, jump, field_cnt,
                                    r & 0xFFFFFFFF, C.byref(s), C.byref(e), C.byref(f)))
    return (s.value, e.value) if f.value else None


def seek_field(index: np.ndarray, data_len: int, record_cnt: int, field_cnt: int, crlf: bool, r: int, fld: int):
    s, e, f = C.c_uint64(), C.c_uint64(), C.c_int()
    _check(lib().oracle_seek_field(index.ctypes.data, index.size, data_len, record_cnt, field_cnt, int(crlf),
                                   r & 0xFFFFFFFF, fld & 0xFFFFFFFF, C.byref(s), C.byref(e), C.byref(f)))
    return (s.value, e.value) if f.value else None


def seek_fields_timed(index: np.ndarray, data_len: int, record_cnt: int, field_cnt: int, crlf: bool,
                      rec: np.ndarray, fld: np.ndarray):
    """CPU baseline of the batched lookup config: returns (checksum, hits)."""
    cs, h = C.c_uint64(), C.c_uint64()
    _check(lib().oracle_seek_fields_timed(index.ctypes.data, index.size, data_len, record_cnt, field_cnt, int(crlf),
                                          rec.ctypes.data, fld.ctypes.data, rec.size, C.byref(cs), C.byref(h)))
    return cs.value, h.value


def boundaries(task_size: int, job_count: int):
    out = (_Boundary * 256)()
    n = lib().oracle_boundaries(task_size, job_count, out)
    if n == 0:
        return None
    return [(out[i].start, out[i].len) for i in range(n)]


def chunks(record_cnt: int, jump: int, num: int):
    out = (_Chunk * 256)()
    n = lib().oracle_chunks(record_cnt, jump, num, out)
    if n == 0:
        return None
    return [dict(id=out[i].id, start=out[i].start, end=out[i].end, record_cnt=out[i].record_cnt)
            for i in range(n)]


# ---- definitions beyond the reference: see csv_oracle.c -----------------------------
def tape_first_bad_slot(data, index: np.ndarray, field_cnt: int, crlf: bool) -> int:
    a = np.ascontiguousarray(np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data)
    index = np.ascontiguousarray(index, dtype=np.uint64)
    return int(lib().oracle_tape_first_bad_slot(a.ctypes.data, a.size, index.ctypes.data, index.size, field_cnt,
                                                int(crlf)))


def field_value(raw: bytes, flags: int) -> bytes:
    buf = np.frombuffer(raw, dtype=np.uint8) if len(raw) else np.zeros(1, dtype=np.uint8)
    out = np.zeros(max(len(raw), 1), dtype=np.uint8)
    n = lib().oracle_field_value(buf.ctypes.data, len(raw), flags, out.ctypes.data)
    return out[:n].tobytes()


def materialize_column(data, index: np.ndarray, record_cnt: int, field_cnt: int, crlf: bool, field_idx: int,
                       first_record: int, nrec: int, flags: int):
    """(offsets[nrec+1], packed bytes): field `field_idx` of records first_record .. +nrec through the
    seek_field restatement (a record that seek_field reports as None contributes an empty value)."""
    a = np.frombuffer(data, dtype=np.uint8) if isinstance(data, (bytes, bytearray)) else np.ascontiguousarray(data, dtype=np.uint8)
    index = np.ascontiguousarray(index, dtype=np.uint64)
    offs = np.zeros(nrec + 1, dtype=np.uint64)
    args = (a.ctypes.data, a.size, index.ctypes.data, index.size, record_cnt, field_cnt, int(crlf), field_idx,
            first_record, nrec, flags, offs.ctypes.data)
    total = lib().oracle_materialize_column(*args, None)
    out = np.zeros(max(int(total), 1), dtype=np.uint8)
    lib().oracle_materialize_column(*args, out.ctypes.data)
    return offs, out[:int(total)].tobytes()


def is_ascii(data) -> bool:
    """reader::is_ascii (src/reader.rs:26-132)."""
    a = np.frombuffer(data, dtype=np.uint8) if isinstance(data, (bytes, bytearray)) else np.ascontiguousarray(data)
    if a.size == 0:
        return True
    return bool(lib().oracle_is_ascii(a.ctypes.data, a.size))


def utf8_valid_up_to(data):
    """core::str::from_utf8(data): None when well-formed, else Utf8Error::valid_up_to().  CPython's strict
    decoder implements the same Unicode well-formedness table and reports the same start position."""
    try:
        bytes(data).decode("utf-8")
        return None
    except UnicodeDecodeError as e:
        return e.start


def blsr(x: int) -> int:
    return lib().oracle_blsr(x)
