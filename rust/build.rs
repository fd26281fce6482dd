// Replacement for the reference build.rs (build.rs:1-13), which asks ec-gpu-gen to emit and
// compile OpenCL/CUDA multiexp source at cargo-build time. The kernels now live in libb200msm.so
// (built by `make -C ark_blst_b200/csrc`, sm_100a only); cargo only has to find and link it.
fn main() {
    #[cfg(feature = "b200")]
    {
        let dir = std::env::var("B200MSM_LIB_DIR").unwrap_or_else(|_| "/usr/local/lib".into());
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-lib=dylib=b200msm");
        println!("cargo:rerun-if-env-changed=B200MSM_LIB_DIR");
    }
}
