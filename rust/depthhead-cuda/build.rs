fn main() {
    // directory holding libdepthhead_cuda.so (built by `python -m depthhead_b200._build`)
    let dir = std::env::var("DEPTHHEAD_CUDA_LIB_DIR").expect("set DEPTHHEAD_CUDA_LIB_DIR");
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=depthhead_cuda");
}
