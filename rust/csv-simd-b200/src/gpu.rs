//! One `csvb200_ctx` per GPU and per thread at a time (include/csvb200.h "Conventions").
use crate::{check, sys, StructureError};
use std::ffi::CStr;
use std::ptr;

pub struct Context {
    raw: *mut sys::csvb200_ctx,
}

// the context is used by one thread at a time; built indexes are immutable and may be queried concurrently
unsafe impl Send for Context {}

impl Context {
    pub fn new(device: i32) -> Result<Self, StructureError> {
        let mut raw = ptr::null_mut();
        check(unsafe { sys::csvb200_ctx_create(device, &mut raw) }, || {
            format!("csvb200_ctx_create({}): no usable CUDA device (there is no CPU fallback)", device)
        })?;
        Ok(Context { raw })
    }
    pub fn raw(&self) -> *mut sys::csvb200_ctx {
        self.raw
    }
    pub fn last_error(&self) -> String {
        unsafe { CStr::from_ptr(sys::csvb200_last_error(self.raw)) }.to_string_lossy().into_owned()
    }
}

impl Drop for Context {
    fn drop(&mut self) {
        unsafe { sys::csvb200_ctx_destroy(self.raw) }
    }
}

/// Every GPU of this process behind one handle (`csvb200_multi_*`).
pub struct Multi {
    raw: *mut sys::csvb200_multi,
}
unsafe impl Send for Multi {}

impl Multi {
    pub fn new(devices: &[i32]) -> Result<Self, StructureError> {
        let mut raw = ptr::null_mut();
        check(unsafe { sys::csvb200_multi_create(devices.as_ptr(), devices.len() as i32, &mut raw) }, || {
            format!("csvb200_multi_create({:?}) failed", devices)
        })?;
        Ok(Multi { raw })
    }
    pub fn raw(&self) -> *mut sys::csvb200_multi {
        self.raw
    }
    pub fn last_error(&self) -> String {
        unsafe { CStr::from_ptr(sys::csvb200_multi_last_error(self.raw)) }.to_string_lossy().into_owned()
    }
}

impl Drop for Multi {
    fn drop(&mut self) {
        unsafe { sys::csvb200_multi_destroy(self.raw) }
    }
}

thread_local! {
    static CTX: Context = Context::new(0).expect("no CUDA device: csv-simd-b200 has no CPU fallback");
}

/// The calling thread's context on device 0.
pub fn with_context<R>(f: impl FnOnce(&Context) -> R) -> R {
    CTX.with(|c| f(c))
}

/// Pins a slice the caller owns for as long as the guard lives (`csvb200_host_register`), so `reader::read` and the
/// lookups DMA it in place instead of staging it through host copies.  ~25-30 ms per GiB: for buffers used more than once
/// (an index `Vec<usize>` that is refilled, an `Mmap` that is indexed and then searched).
pub struct Pinned<'a, T> {
    slice: &'a [T],
}

impl<'a, T> Pinned<'a, T> {
    pub fn new(slice: &'a [T], read_only: bool) -> Result<Self, StructureError> {
        let bytes = std::mem::size_of_val(slice);
        check(
            unsafe { sys::csvb200_host_register(slice.as_ptr() as *mut std::ffi::c_void, bytes, read_only as i32) },
            || "csvb200_host_register: the range is already registered or cannot be pinned".to_string(),
        )?;
        Ok(Pinned { slice })
    }
}

impl<'a, T> Drop for Pinned<'a, T> {
    fn drop(&mut self) {
        unsafe { sys::csvb200_host_unregister(self.slice.as_ptr() as *mut std::ffi::c_void) };
    }
}
