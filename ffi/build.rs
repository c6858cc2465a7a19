// Links libsrt.so (and, with the `nccl` feature, libsrt_nccl.so).  SRT_LIB_DIR = the directory that holds them
// (spectral_raytracer_b200/ of this repository after `python -c "import __graft_entry__ as g; g.build()"`).
fn main() {
    println!("cargo:rerun-if-env-changed=SRT_LIB_DIR");
    if let Ok(dir) = std::env::var("SRT_LIB_DIR") {
        println!("cargo:rustc-link-search=native={}", dir);
        println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
    }
    println!("cargo:rustc-link-lib=dylib=srt");
    if std::env::var("CARGO_FEATURE_NCCL").is_ok() {
        println!("cargo:rustc-link-lib=dylib=srt_nccl");
    }
}
