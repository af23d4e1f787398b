// Builds libsylt2d_b200.so from ../sylt2d_b200/csrc with nvcc for sm_100a and links it.
// -fmad=false / -ffp-contract=off are part of the numerical contract (the reference never fuses a*b+c).
use std::{env, path::PathBuf, process::Command};

fn main() {
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let csrc = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../sylt2d_b200/csrc");
    let so = out.join("libsylt2d_b200.so");
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".into());
    let status = Command::new(nvcc)
        .current_dir(&csrc)
        .args(["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false"])
        .args(["-Xcompiler", "-fPIC,-ffp-contract=off", "-shared", "-o"])
        .arg(&so)
        .args(["world.cu", "batch.cu"])
        .status()
        .expect("nvcc not found: the sylt2d_b200 backend has no CPU fallback");
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=sylt2d_b200");
    println!("cargo:rustc-link-lib=dylib=cudart");
    for f in ["world.cu", "batch.cu", "narrow.cuh", "solver.cuh", "sylt_math.cuh", "sylt_sincos.h", "common.h"] {
        println!("cargo:rerun-if-changed={}", csrc.join(f).display());
    }
}
