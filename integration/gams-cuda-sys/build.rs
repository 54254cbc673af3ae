// Links libgams_cuda.so.  GAMS_CUDA_LIB_DIR points at the directory holding it
// (gams_b200/ in this repo after `make -C gams_b200/csrc`).
fn main() {
    if let Ok(dir) = std::env::var("GAMS_CUDA_LIB_DIR") {
        println!("cargo:rustc-link-search=native={}", dir);
        println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
    }
    println!("cargo:rustc-link-lib=dylib=gams_cuda");
    println!("cargo:rerun-if-env-changed=GAMS_CUDA_LIB_DIR");
}
