// Link against the in-tree shared library: SANDENGINE_B200_LIB_DIR must point at the directory that holds
// libsandengine_b200.so (the `sandengine_b200/` package directory of this repository after `build()`).
fn main() {
    if let Ok(dir) = std::env::var("SANDENGINE_B200_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    }
    println!("cargo:rustc-link-lib=dylib=sandengine_b200");
    println!("cargo:rerun-if-env-changed=SANDENGINE_B200_LIB_DIR");
}
