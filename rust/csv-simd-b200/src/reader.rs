//! Replacement body of csv-simd's `reader::read` (src/reader.rs:150-306).
//!
//! The signature, the returned type (`Vec<usize>` behind `StructureIndex(Vec<CodeUnitPos>)`, src/stage1.rs:61,67)
//! and the panic on inputs shorter than 64 bytes (src/reader.rs:220-229, src/avx/stage1.rs:45-48) are kept; the
//! index is bit-exact: `[0] ++ [ i | b[i] in {',', CR, LF} and an even number of '"' before i ]`.
use crate::gpu::{with_context, Multi};
use crate::{check, sys, StructureError};
use memmap::Mmap;

/// `csv -> memory index` on one GPU: H2D of the mapped bytes, the fused index kernel and D2H of the finished
/// segments overlap chunk by chunk (`csvb200_index_build_to_host`), so the call costs about max(upload, download).
pub fn read(memmap: &Mmap) -> Vec<usize> {
    assert!(memmap.len() >= 64, "reader::read: the reference panics on inputs shorter than 64 bytes");
    read_bytes(&memmap[..]).expect("csvb200_index_build_to_host")
}

pub fn read_bytes(bytes: &[u8]) -> Result<Vec<usize>, StructureError> {
    with_context(|gpu| {
        // entries <= bytes + 1; a third of the input is the library's own first guess, the call reports the exact
        // size when the guess is too small and we go again with it
        let mut cap = bytes.len() / 3 + 4096;
        loop {
            let mut acc: Vec<usize> = Vec::with_capacity(cap);
            let mut len: usize = 0;
            let rc = unsafe {
                sys::csvb200_index_build_to_host(gpu.raw(), bytes.as_ptr(), bytes.len(), acc.as_mut_ptr() as *mut u64, cap,
                                                 &mut len)
            };
            if rc == sys::CSVB200_ERR_CAPACITY && len > cap {
                cap = len;
                continue;
            }
            check(rc, || gpu.last_error())?;
            unsafe { acc.set_len(len) };
            return Ok(acc);
        }
    })
}

/// The same over all listed GPUs of this process: the slice is cut into `devices.len()` contiguous byte ranges at
/// arbitrary offsets, every GPU indexes its range under a predicted quote carry, the 32-byte rows cross NVLink from
/// inside the kernels, a shard whose guess was wrong is re-indexed, and the segments land at their final places in
/// ONE `Vec<usize>` -- `reader::read` at 8 GPUs (src/lib.rs:61-74 keeps calling it with one `&Mmap`).
pub fn read_multi(multi: &Multi, bytes: &[u8]) -> Result<Vec<usize>, StructureError> {
    let mut cap = bytes.len() / 3 + 4096;
    loop {
        let mut acc: Vec<usize> = Vec::with_capacity(cap);
        let mut len: usize = 0;
        let rc = unsafe {
            sys::csvb200_multi_index_build_to_host(multi.raw(), bytes.as_ptr(), bytes.len(), std::ptr::null(),
                                                   acc.as_mut_ptr() as *mut u64, cap, &mut len)
        };
        if rc == sys::CSVB200_ERR_CAPACITY && len > cap {
            cap = len;
            continue;
        }
        check(rc, || multi.last_error())?;
        unsafe { acc.set_len(len) };
        return Ok(acc);
    }
}
