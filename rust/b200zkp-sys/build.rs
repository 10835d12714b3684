// NOT BUILT IN THIS ENVIRONMENT — see INTEGRATION.md.
// Compiles intmax_zkp_core_b200/csrc/b200zkp.cu for sm_100a with nvcc through the `cc` crate.
fn main() {
    let root = std::path::Path::new(env!("CARGO_MANIFEST_DIR")).join("../..");
    let src = root.join("intmax_zkp_core_b200/csrc/b200zkp.cu");
    println!("cargo:rerun-if-changed={}", src.display());
    cc::Build::new()
        .cuda(true)
        .cudart("shared")
        .flag("-gencode").flag("arch=compute_100a,code=sm_100a")
        .flag("-O3").flag("-lineinfo").flag("-std=c++17")
        .include(root.join("include"))
        .file(src)
        .compile("b200zkp");
    // libnccl.so.2 is loaded at run time (dlopen) by the multi-GPU entry points only
    println!("cargo:rustc-link-lib=dl");
}
