// Links libkopek_cuda.so.  KOPEK_CUDA_LIB_DIR points at the directory that holds it
// (kopek_b200/lib in this repository after `python -m kopek_b200.build`).
fn main() {
    let dir = std::env::var("KOPEK_CUDA_LIB_DIR").unwrap_or_else(|_| "../../kopek_b200/lib".to_string());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=kopek_cuda");
    println!("cargo:rerun-if-env-changed=KOPEK_CUDA_LIB_DIR");
}
