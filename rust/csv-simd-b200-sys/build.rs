// Links libcsvb200.so.  CSVB200_LIB_DIR points at the directory that holds it (csv_simd_b200/ in this repository,
// where `python csv_simd_b200/build.py` puts it); the rpath is set so that binaries find it without LD_LIBRARY_PATH.
use std::env;
use std::path::PathBuf;

fn main() {
    let dir = env::var("CSVB200_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| {
        PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../../csv_simd_b200")
    });
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=csvb200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
    println!("cargo:rerun-if-env-changed=CSVB200_LIB_DIR");
    println!("cargo:rerun-if-changed=../../include/csvb200.h");
}
