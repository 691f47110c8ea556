// Link against the prebuilt libsearchlite_gpu.so (built by `python -m searchlite_b200.build`, nvcc, sm_100a).
// SEARCHLITE_GPU_LIB_DIR names the directory that holds it; the default is this repository's in-tree build output.
use std::env;
use std::path::PathBuf;

fn main() {
    let dir = env::var("SEARCHLITE_GPU_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| {
        PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../../searchlite_b200/lib")
    });
    println!("cargo:rerun-if-env-changed=SEARCHLITE_GPU_LIB_DIR");
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=searchlite_gpu");
    // let test binaries find the library without LD_LIBRARY_PATH
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
}
