// build.rs — compiles the CUDA library with nvcc for sm_100a and links it.
// UNCOMPILED in this image (no Rust toolchain); mirrors dark_b200/csrc/Makefile.
use std::env;
use std::path::PathBuf;
use std::process::Command;

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..");
    let csrc = root.join("dark_b200/csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".to_string());
    let lib = out.join("libdark_bwt.so");
    let status = Command::new(&nvcc)
        .args(&["-O3", "-std=c++17", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
                "-Xcompiler", "-fPIC", "-shared", "-o"])
        .arg(&lib)
        .arg(csrc.join("dark_bwt.cu"))
        .arg(csrc.join("synth.cpp"))
        .status()
        .expect("nvcc not found: the forward BWT has no CPU fallback");
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=dark_bwt");
    for f in &["dark_bwt.cu", "synth.cpp", "common.cuh", "radix_sort.cuh", "suffix_kernels.cuh"] {
        println!("cargo:rerun-if-changed={}", csrc.join(f).display());
    }
    println!("cargo:rerun-if-changed={}", root.join("include/dark_bwt.h").display());
}
