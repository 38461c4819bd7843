"""Host-side mirror of the reference's public API for the hot path (same names, argument meaning
and error behaviour), sitting on top of the C ABI:

    create(filename) -> Tape                      src/lib.rs:61-74
    reader.read(bytes) -> StructureIndex          src/reader.rs:150-306   (GPU: csvb200_index_build)
    Header.new(bytes)                             src/tape.rs:226-273     (host scalar scan of line 1)
    TapeCore.create / init, Tape.from_core        src/tape.rs:303-347, 83-94
    Tape.chunks(num), boundaries(task, jobs)      src/tape.rs:95-140, 385-428
    RecordSource.seek_record / seek_field         src/record_source.rs:70-140

Only O(first line) / O(1) metadata work happens here; the index build and the batched lookups run
on the GPU through libcsvb200.  There is no CPU implementation of the index build in this package.
"""
from __future__ import annotations

import mmap as _mmap
import os
from dataclasses import dataclass
from enum import Enum
from typing import List, Optional

import numpy as np

from . import api
from .errors import InvalidCsvFormat, InvalidState, Io, ReferencePanic

_default_ctx: Optional[api.Context] = None


def default_context() -> api.Context:
    global _default_ctx
    if _default_ctx is None or _default_ctx._h is None:
        _default_ctx = api.Context(int(os.environ.get("LOCAL_RANK", "0")))
    return _default_ctx


class NewLine(Enum):
    """src/stage1.rs:472-480"""
    CRLF = "CRLF"
    LF = "LF"


class Mmap:
    """Stand-in for memmap::Mmap (src/lib.rs:64-65): a read-only view of the file's bytes."""

    def __init__(self, path: Optional[str] = None, data=None):
        self._f = None
        self._mm = None
        if path is not None:
            try:
                self._f = open(path, "rb")
                size = os.fstat(self._f.fileno()).st_size
                if size == 0:
                    self._arr = np.zeros(0, dtype=np.uint8)
                else:
                    self._mm = _mmap.mmap(self._f.fileno(), 0, access=_mmap.ACCESS_READ)
                    self._arr = np.frombuffer(self._mm, dtype=np.uint8)
            except OSError as e:  # File::open(...)? / Mmap::map(...)? -> StructureError::Io
                raise Io(e.errno, e.strerror, path) from e
        else:
            self._arr = api._as_u8(data)

    @staticmethod
    def map(path: str) -> "Mmap":
        return Mmap(path=path)

    def __len__(self):
        return int(self._arr.size)

    @property
    def array(self) -> np.ndarray:
        return self._arr

    def slice(self, start: int, end: int) -> bytes:
        return self._arr[start:end].tobytes()


@dataclass
class Header:
    """src/tape.rs:217-223"""
    header: List[str]
    new_line: NewLine
    field_cnt: int
    delimiter: int
    record_offset: int

    @staticmethod
    def new(memmap: Mmap) -> "Header":
        """Header::new (src/tape.rs:226-273): first CR/LF ends the header (:228-232); CRLF iff the byte
        after it is LF (:235-238); leading EF/BB/BF bytes are skipped (:241-249); names are
        split(",") + trim, NOT quote aware (:259-262)."""
        b = memmap.array
        n = b.size
        # only the first line is scanned; grow the window until it contains a line end
        end = n
        w = 4096
        while True:
            win = b[:min(w, n)]
            hit = np.flatnonzero((win == 0x0D) | (win == 0x0A))
            if hit.size:
                end = int(hit[0])
                break
            if w >= n:
                break
            w *= 16
        if end + 1 >= n:
            raise ReferencePanic("Header::new indexes memmap[header_end_idx + 1] out of bounds (src/tape.rs:236)")
        new_line = NewLine.CRLF if b[end + 1] == 0x0A else NewLine.LF
        start = 0
        while start < n and b[start] in (0xEF, 0xBB, 0xBF):
            start += 1
        if start > end:
            raise ReferencePanic("Header::new slices memmap[start..end] with start > end (src/tape.rs:253)")
        text = b[start:end].tobytes().decode("utf-8", "replace")
        names = [s.strip() for s in text.split(",")]
        return Header(names, new_line, len(names), 0x2C, end)


class reader:
    """src/reader.rs"""

    @staticmethod
    def read(memmap, ctx: Optional[api.Context] = None, keep_bytes: bool = True) -> api.StructureIndex:
        """reader::read (src/reader.rs:150-306).  Inputs shorter than 64 bytes make the reference
        panic (:220-229 + src/avx/stage1.rs:45-48); that is mirrored as ReferencePanic."""
        ctx = ctx or default_context()
        arr = memmap.array if isinstance(memmap, Mmap) else api._as_u8(memmap)
        flags = api.BUILD_STRICT_MIN64 | (api.BUILD_KEEP_BYTES if keep_bytes else 0)
        return ctx.index_build(arr, flags)


@dataclass
class Boundary:
    """src/tape.rs:281-284"""
    start: int
    len: int


def boundaries(task_size: int, job_count: int) -> Optional[List[Boundary]]:
    """boundaries(task_size: u32, job_count: u8) (src/tape.rs:385-428)."""
    task_size &= 0xFFFFFFFF
    job_count &= 0xFF
    if task_size == 0 or job_count == 0:
        return None
    if task_size < job_count:
        return [Boundary(0, task_size)]
    job_size, remainder = divmod(task_size, job_count)
    out, acc_end, share = [], 0, 1
    for i in range(job_count):
        if share == 1 and i >= (remainder & 0xFF):
            share = 0
        out.append(Boundary(acc_end, job_size + share))
        acc_end += job_size + share
    return out


@dataclass
class Chunk:
    """src/tape.rs:13-19"""
    id: int
    start: int
    end: int
    record_cnt: int
    index: api.StructureIndex


class RecordSource:
    """trait RecordSource (src/record_source.rs:68-147): scalar seeks read two entries of the
    host copy of the index; the *_batch variants run the K4 gather kernel."""

    # accessors supplied by the implementor
    def record_cnt(self) -> Optional[int]:
        raise NotImplementedError

    def index(self) -> api.StructureIndex:
        raise NotImplementedError

    def record_jump_size(self) -> int:
        raise NotImplementedError

    def field_cnt(self) -> int:
        raise NotImplementedError

    def new_line_tag(self) -> NewLine:
        raise NotImplementedError

    def data_bytes(self) -> Mmap:
        raise NotImplementedError

    def seek_record(self, record_idx: int) -> Optional[str]:
        """src/record_source.rs:70-102"""
        rc = self.record_cnt()
        if rc is None:
            raise InvalidState()
        record_idx &= 0xFFFFFFFF
        if ((record_idx + 1) & 0xFFFFFFFF) >= rc:
            return None
        field_cnt = self.field_cnt()
        idx_start = (((record_idx + 1) & 0xFFFFFFFF) * (self.record_jump_size() & 0xFFFFFFFF)) & 0xFFFFFFFF
        host = self.index().host()
        if idx_start + field_cnt >= host.size:
            raise ReferencePanic("index out of bounds (src/record_source.rs:94-95)")
        s, e = int(host[idx_start]) + 1, int(host[idx_start + field_cnt])
        return self._slice(s, e)

    def seek_field(self, record_idx: int, field_idx: int) -> Optional[str]:
        """src/record_source.rs:104-140 (the five unconditional println! are not reproduced)"""
        rc = self.record_cnt()
        if rc is None:
            raise InvalidState()
        record_idx &= 0xFFFFFFFF
        field_idx &= 0xFFFFFFFF
        if ((record_idx + 1) & 0xFFFFFFFF) >= rc:
            return None
        field_cnt = self.field_cnt()
        if field_idx >= field_cnt:
            return None
        row_size = field_cnt + 1 if self.new_line_tag() is NewLine.CRLF else field_cnt
        idx_start = (((record_idx + 1) & 0xFFFFFFFF) * row_size + field_idx) & 0xFFFFFFFF
        host = self.index().host()
        if idx_start + 1 >= host.size:
            raise ReferencePanic("index out of bounds (src/record_source.rs:132-133)")
        s, e = int(host[idx_start]) + 1, int(host[idx_start + 1])
        return self._slice(s, e)

    def _slice(self, s: int, e: int) -> str:
        data = self.data_bytes()
        if s > e or e > len(data):
            raise ReferencePanic("slice index out of range")
        return data.slice(s, e).decode("utf-8", "replace")

    # batched (GPU) forms: ranges[(start, end)], (UINT64_MAX, UINT64_MAX) = None
    def seek_fields_batch(self, rec, fld) -> np.ndarray:
        if self.record_cnt() is None:
            raise InvalidState()
        return self.index().seek_fields(rec, fld)

    def seek_records_batch(self, rec) -> np.ndarray:
        if self.record_cnt() is None:
            raise InvalidState()
        return self.index().seek_records(rec)


class TapeCore(RecordSource):
    """src/tape.rs:185-212, 301-352"""

    def __init__(self, memmap: Mmap, index: api.StructureIndex, header: Header):
        self.header_ = header
        self.index_ = index
        self.memmap = memmap
        self.first_record_idx = None
        self.record_cnt_: Optional[int] = None
        self.record_jump_size_: Optional[int] = None

    @staticmethod
    def create(memmap: Mmap, index: api.StructureIndex, header: Header) -> "TapeCore":
        return TapeCore(memmap, index, header)

    def init(self):
        """TapeCore::init (src/tape.rs:315-347) via csvb200_tape_init."""
        crlf = self.header_.new_line is NewLine.CRLF
        try:
            rc, jump = self.index_.tape_init(self.header_.field_cnt, crlf)
        except InvalidCsvFormat:
            # the reference sets both fields before it returns the error (:318-325, 342-344)
            n = len(self.index_)
            jump = self.header_.field_cnt + (1 if crlf else 0)
            self.record_jump_size_ = jump
            self.record_cnt_ = ((n - 1) // jump) & 0xFFFFFFFF
            raise
        self.record_jump_size_ = jump
        self.record_cnt_ = rc

    def header(self) -> List[str]:
        return self.header_.header

    def record_cnt(self):
        return self.record_cnt_

    def index(self):
        return self.index_

    def record_jump_size(self) -> int:
        if self.record_jump_size_ is None:
            raise InvalidState()
        return self.record_jump_size_

    def field_cnt(self) -> int:
        return self.header_.field_cnt

    def new_line_tag(self) -> NewLine:
        return self.header_.new_line

    def data_bytes(self) -> Mmap:
        return self.memmap


class Tape(RecordSource):
    """src/tape.rs:74-174"""

    def __init__(self, header: Header, record_cnt: int, record_jump_size: int, bytes_: Mmap,
                 index: api.StructureIndex):
        self.header_ = header
        self.record_cnt_ = record_cnt
        self.record_jump_size_ = record_jump_size
        self.bytes_ = bytes_
        self.index_ = index

    @staticmethod
    def from_core(core: TapeCore) -> "Tape":
        core.init()
        return Tape(core.header_, core.record_cnt_, core.record_jump_size_, core.memmap, core.index_)

    def chunks(self, num: int) -> List[Chunk]:
        """Tape::chunks(num: u8) (src/tape.rs:95-140)."""
        bs = boundaries(self.record_cnt_, num)
        if bs is None:
            raise InvalidState()
        j = self.record_jump_size_
        out = [Chunk(i & 0xFF, b.start * j, (b.start + b.len) * j, b.len & 0xFFFFFFFF, self.index_)
               for i, b in enumerate(bs)]
        c0 = out[0]
        out[0] = Chunk(c0.id, j, c0.end, (c0.record_cnt - 1) & 0xFFFFFFFF, c0.index)
        return out

    def index(self):
        return self.index_

    def bytes(self) -> Mmap:
        return self.bytes_

    def header(self) -> List[str]:
        return self.header_.header

    def record_cnt(self):
        return self.record_cnt_

    def record_jump_size(self) -> int:
        return self.record_jump_size_

    def field_cnt(self) -> int:
        return self.header_.field_cnt

    def new_line_tag(self) -> NewLine:
        return self.header_.new_line

    def data_bytes(self) -> Mmap:
        return self.bytes_


def create(filename: str, ctx: Optional[api.Context] = None) -> Tape:
    """csv_simd::create (src/lib.rs:61-74): open -> mmap -> Header::new -> reader::read ->
    TapeCore::create -> Tape::from_core."""
    memmap = Mmap.map(filename)
    header = Header.new(memmap)
    index = reader.read(memmap, ctx)
    core = TapeCore.create(memmap, index, header)
    return Tape.from_core(core)
