// Links libmodppl_b200.so (built by `make` at the repository root into modppl_b200/lib/).
// MODPPL_B200_LIB_DIR overrides the search directory for an installed copy.
fn main() {
    let dir = std::env::var("MODPPL_B200_LIB_DIR").unwrap_or_else(|_| {
        let here = std::path::PathBuf::from(std::env::var("CARGO_MANIFEST_DIR").unwrap());
        here.join("../../modppl_b200/lib").to_string_lossy().into_owned()
    });
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=modppl_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=MODPPL_B200_LIB_DIR");
}
