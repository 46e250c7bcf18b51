// Links libbellman_b200.so, built in-tree by `python -c 'import __graft_entry__ as g; g.build()'`.
// BELLMAN_B200_LIB_DIR overrides the search path (default: ../../bellman_b200 relative to this crate).
use std::env;
use std::path::PathBuf;

fn main() {
    let dir = env::var("BELLMAN_B200_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| {
        PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../../bellman_b200")
    });
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=bellman_b200");
    println!("cargo:rerun-if-env-changed=BELLMAN_B200_LIB_DIR");
}
