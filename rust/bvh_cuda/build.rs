// Links the prebuilt libbvh_cuda.so (built by `make -C voidin_b200/csrc`); set BVH_CUDA_LIB_DIR to its directory.
fn main() {
    let dir = std::env::var("BVH_CUDA_LIB_DIR").unwrap_or_else(|_| "../../voidin_b200".into());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=bvh_cuda");
    println!("cargo:rerun-if-env-changed=BVH_CUDA_LIB_DIR");
}
