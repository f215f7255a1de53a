// Links libflacb200.so (built by `python -m flac_codec_b200.build`); FLACB200_LIB_DIR points at flac_codec_b200/.
fn main() {
    let dir = std::env::var("FLACB200_LIB_DIR").unwrap_or_else(|_| "../../flac_codec_b200".into());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=flacb200");
    println!("cargo:rerun-if-env-changed=FLACB200_LIB_DIR");
}
