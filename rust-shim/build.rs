// Links libcircom_witnesscalc.so (built by `python circom-witnesscalc_b200/build.py`).  The reference's build.rs
// (build.rs:9-30) runs bindgen over include/graph_witness.h and prost over protos/messages.proto; neither is needed
// here: the few C types are declared by hand in src/ffi.rs, and the graph file is parsed by the library itself.
use std::env;
use std::path::PathBuf;

fn main() {
    let dir = env::var("GW_B200_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| {
        PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../circom-witnesscalc_b200/lib")
    });
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=circom_witnesscalc");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
    println!("cargo:rerun-if-env-changed=GW_B200_LIB_DIR");
    println!("cargo:rerun-if-changed=../include/graph_witness.h");
}
