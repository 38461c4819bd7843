"""Error types mirroring the reference's StructureError (src/error.rs:9-21) plus the GPU-side ones."""
from __future__ import annotations


class StructureError(Exception):
    """Base class: csv_simd::StructureError."""


class Io(StructureError, OSError):
    """StructureError::Io (src/error.rs:11-12)."""


class MissingValue(StructureError):
    """StructureError::MissingValue (src/error.rs:15-16)."""

    def __str__(self):
        return "Missing a value"


class InvalidState(StructureError):
    """StructureError::InvalidState (src/error.rs:17-18)."""

    def __str__(self):
        return "Invalid state"


class InvalidCsvFormat(StructureError):
    """StructureError::InvalidCsvFormat (src/error.rs:19-20)."""

    def __str__(self):
        return "Unsupported csv structure: likely variable number of fields"


class ReferencePanic(StructureError):
    """Inputs on which the reference panics (n < 64 bytes: src/reader.rs:220-229,
    src/avx/stage1.rs:45-48; out-of-bounds index slot: src/record_source.rs:94-95,132-133)."""


class GpuError(StructureError):
    """CUDA / device-side failure in libcsvb200 (no CPU fallback exists)."""


OK = 0
_BY_CODE = {
    1: ValueError,
    2: InvalidState,
    3: InvalidCsvFormat,
    4: MissingValue,
    5: Io,
    6: GpuError,
    7: MemoryError,
    8: ReferencePanic,
    9: BufferError,
    10: ReferencePanic,
    11: GpuError,
}


def raise_for(code: int, detail: str = ""):
    if code == OK:
        return
    exc = _BY_CODE.get(code, GpuError)
    if exc in (InvalidState, InvalidCsvFormat, MissingValue):
        raise exc()
    raise exc(detail or f"libcsvb200 status {code}")
