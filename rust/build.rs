// Links libcute_nucleotides_cuda.so (built by `make -C cute_nucleotides_b200/csrc`, nvcc sm_100a).
// CN_CUDA_LIB_DIR overrides the directory that holds the shared library.
use std::env;
use std::path::PathBuf;

fn main() {
    let dir = env::var("CN_CUDA_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| {
        PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../cute_nucleotides_b200")
    });
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=cute_nucleotides_cuda");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
    println!("cargo:rerun-if-env-changed=CN_CUDA_LIB_DIR");
}
