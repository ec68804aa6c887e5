// Points the linker at the in-tree libb200gs.so: B200GS_LIB_DIR=<repo>/wgpu-3dgs-viewer-app_b200 cargo build
fn main() {
    let dir = std::env::var("B200GS_LIB_DIR").expect("set B200GS_LIB_DIR to the directory that holds libb200gs.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=b200gs");
    println!("cargo:rerun-if-env-changed=B200GS_LIB_DIR");
}
