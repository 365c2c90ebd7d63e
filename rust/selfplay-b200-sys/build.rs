// Links against libselfplay_b200.so built by `make -C self-play-ai_b200/csrc`.
// SELFPLAY_B200_LIB_DIR overrides the default in-tree location.
fn main() {
    let dir = std::env::var("SELFPLAY_B200_LIB_DIR").unwrap_or_else(|_| {
        let manifest = std::env::var("CARGO_MANIFEST_DIR").unwrap();
        format!("{}/../../self-play-ai_b200", manifest)
    });
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=selfplay_b200");
    println!("cargo:rerun-if-env-changed=SELFPLAY_B200_LIB_DIR");
}
