//! Points the linker at libstrolle_b200.so.  `STROLLE_B200_LIB_DIR` = the directory that holds it (the repository builds it in-tree:
//! `python -m strolle_b200.build` -> strolle_b200/_lib/); falls back to that in-tree path relative to this crate.
use std::env;
use std::path::PathBuf;

fn main() {
    println!("cargo:rerun-if-env-changed=STROLLE_B200_LIB_DIR");
    let dir = env::var_os("STROLLE_B200_LIB_DIR")
        .map(PathBuf::from)
        .unwrap_or_else(|| PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../../strolle_b200/_lib"));
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=strolle_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
}
