"""csv_simd_b200 -- B200-native (sm_100a) drop-in for csv-simd's hot path:
csv -> in-memory structural index -> (record #) -> record -> (record, field #) -> field.

Compute lives in csrc/ (hand-written CUDA behind the C ABI in include/csvb200.h); this package is
the ctypes binding plus the host-side mirror of the reference's public API.  No CPU fallback.
"""
from .api import (BUILD_DEFAULT, BUILD_KEEP_BYTES, BUILD_STRICT_MIN64, BUILD_VALIDATE, FIELD_RAW, FIELD_TRIM,  # noqa: F401
                  FIELD_UNQUOTE, Context, Exchange, Multi, StructureIndex, host_registered)
from . import errors  # noqa: F401
from .errors import (GpuError, InvalidCsvFormat, InvalidState, Io, MissingValue, ReferencePanic,  # noqa: F401
                     StructureError)
from .tape import (Boundary, Chunk, Header, Mmap, NewLine, RecordSource, Tape, TapeCore, boundaries,  # noqa: F401
                   create, default_context, reader)

__all__ = [
    "Context", "StructureIndex", "Exchange", "Multi", "create", "reader", "Header", "Tape", "TapeCore", "RecordSource", "Mmap",
    "NewLine", "Boundary", "Chunk", "boundaries", "StructureError", "Io", "MissingValue", "InvalidState",
    "InvalidCsvFormat", "ReferencePanic", "GpuError", "BUILD_DEFAULT", "BUILD_KEEP_BYTES", "BUILD_STRICT_MIN64", "BUILD_VALIDATE",
    "FIELD_RAW", "FIELD_UNQUOTE", "FIELD_TRIM",
    "default_context", "host_registered",
]
