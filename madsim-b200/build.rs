// Link against libmsim_b200.so.  MSIM_B200_LIB_DIR points at the directory that holds it
// (default: <repo>/madsim_b200, where `python -c 'import __graft_entry__ as g; g.build()'` puts it).
fn main() {
    let dir = std::env::var("MSIM_B200_LIB_DIR")
        .unwrap_or_else(|_| format!("{}/../madsim_b200", env!("CARGO_MANIFEST_DIR")));
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=msim_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=MSIM_B200_LIB_DIR");
    println!("cargo:rerun-if-changed=../include/msim_b200.h");
}
