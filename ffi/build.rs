// SOURCE ONLY -- compile-unverified (no Rust toolchain in the build image).
// Builds the sm_100a kernels with nvcc through the repo's own build script and links the C ABI.
use std::process::Command;
fn main() {
    let root = std::path::Path::new(env!("CARGO_MANIFEST_DIR")).parent().unwrap().to_path_buf();
    let pkg = root.join("zk-dcap-verifier_b200");
    let status = Command::new("python3").arg(pkg.join("build.py")).status().expect("python + nvcc are required");
    assert!(status.success(), "nvcc build of libzkdcap_b200.so failed");
    println!("cargo:rustc-link-search=native={}", pkg.display());
    println!("cargo:rustc-link-lib=dylib=zkdcap_b200");
    println!("cargo:rerun-if-changed={}", pkg.join("csrc").display());
    println!("cargo:rerun-if-changed={}", root.join("include/zkdcap_b200.h").display());
}
