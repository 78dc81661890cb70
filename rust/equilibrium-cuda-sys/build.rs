// Builds libequilibrium_cuda.so with nvcc for sm_100a and links it.
// NOTE: written for a machine that has cargo + nvcc; the build image of this repo has no
// Rust toolchain, so this file is reviewed by eye and mirrored by equilibrium_b200/build.py.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let so = out.join("libequilibrium_cuda.so");
    let src = root.join("equilibrium_b200/csrc/eq_api.cu");
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".into());
    let status = Command::new(nvcc)
        .args(["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
               "-fmad=false", "-Xcompiler", "-fPIC", "-shared", "-o"])
        .arg(&so)
        .arg(&src)
        .status()
        .expect("nvcc not found");
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=equilibrium_cuda");
    println!("cargo:rerun-if-changed={}", root.join("equilibrium_b200/csrc").display());
    println!("cargo:rerun-if-changed={}", root.join("include/equilibrium_cuda.h").display());
}
