// Links the prebuilt shared library; it is built by `python -m candle_birefnet_b200.build` (nvcc, sm_100a).
fn main() {
    let dir = std::env::var("BIREFNET_B200_LIB_DIR")
        .expect("set BIREFNET_B200_LIB_DIR to the directory that holds libbirefnet_b200.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=birefnet_b200");
    println!("cargo:rerun-if-env-changed=BIREFNET_B200_LIB_DIR");
}
