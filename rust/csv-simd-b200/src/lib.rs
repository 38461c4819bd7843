//! Safe layer over `csv-simd-b200-sys`: what a csv-simd maintainer drops into the crate.
//!
//! * [`gpu::Context`]  one `csvb200_ctx` per GPU, created lazily (`gpu::context()`).
//! * [`reader::read`]  the body that replaces `reader::read` (src/reader.rs:150-306): same signature, the 64-byte
//!   SimdInput loop / `structure()` / `crush_set_bits()` are gone; host bytes go up, the index comes down, uploads and
//!   downloads overlapped chunk by chunk (`csvb200_index_build_to_host`).
//! * [`reader::read_multi`]  the same over every GPU of the box: ONE byte slice in, ONE `Vec<usize>` out
//!   (`csvb200_multi_index_build_to_host`), the file cut "without first knowing record breaks" (README.md:24).
//! * [`record_source::BatchedSeeks`]  `seek_fields` / `seek_records` for slices of queries (K4 gather kernel) next to
//!   the unchanged scalar `seek_record` / `seek_field` (src/record_source.rs:70-140).
//!
//! There is no Rust toolchain in the image this repository is built in: these files are source only and are kept
//! in sync with `include/csvb200.h` by `tools/gen_rust_sys.py` (checked by `tests/test_abi.py`).
pub mod gpu;
pub mod reader;
pub mod record_source;

pub use csv_simd_b200_sys as sys;

/// csv_simd::StructureError (src/error.rs:9-21) as seen from the C status codes.
#[derive(Debug)]
pub enum StructureError {
    Io(String),
    MissingValue,
    InvalidState,
    InvalidCsvFormat,
    /// CUDA / exchange / out-of-memory failures of the device path (there is no CPU fallback)
    Gpu(i32, String),
}

pub(crate) fn check(rc: std::os::raw::c_int, detail: impl FnOnce() -> String) -> Result<(), StructureError> {
    match rc {
        sys::CSVB200_OK => Ok(()),
        sys::CSVB200_ERR_INVALID_STATE => Err(StructureError::InvalidState),
        sys::CSVB200_ERR_INVALID_CSV_FORMAT => Err(StructureError::InvalidCsvFormat),
        sys::CSVB200_ERR_MISSING_VALUE => Err(StructureError::MissingValue),
        sys::CSVB200_ERR_IO => Err(StructureError::Io(detail())),
        // the reference panics on these inputs (assert!(load < 4), Vec bounds check): so does the binding
        sys::CSVB200_ERR_INPUT_TOO_SMALL | sys::CSVB200_ERR_OUT_OF_BOUNDS => panic!("{}", detail()),
        other => Err(StructureError::Gpu(other, detail())),
    }
}
