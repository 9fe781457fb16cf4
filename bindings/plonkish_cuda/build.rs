// Builds libplonkish_cuda.so for sm_100a with nvcc (or links a prebuilt one).
//
//   PLONKISH_CUDA_LIB_DIR=/path/to/dir   link the prebuilt library in that directory
//   PLONKISH_CUDA_SRC=/path/to/repo      repo root holding plonkish_b200/csrc and include/
use std::{env, path::PathBuf, process::Command};

fn main() {
    println!("cargo:rerun-if-env-changed=PLONKISH_CUDA_LIB_DIR");
    println!("cargo:rerun-if-env-changed=PLONKISH_CUDA_SRC");
    if let Ok(dir) = env::var("PLONKISH_CUDA_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-lib=dylib=plonkish_cuda");
        return;
    }
    let src = PathBuf::from(env::var("PLONKISH_CUDA_SRC").expect("set PLONKISH_CUDA_SRC or PLONKISH_CUDA_LIB_DIR"));
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let api = src.join("plonkish_b200/csrc/api.cu");
    let host_copy = src.join("plonkish_b200/csrc/host_copy.cpp");
    for f in [
        "api.cu", "host_copy.cpp", "fq.cuh", "g1.cuh", "msm_kernels.cuh", "poly_kernels.cuh", "sumcheck_kernels.cuh", "lookup_kernels.cuh",
        "dpfq.cuh",
    ] {
        println!("cargo:rerun-if-changed={}", src.join("plonkish_b200/csrc").join(f).display());
    }
    let lib = out.join("libplonkish_cuda.so");
    let status = Command::new(env::var("NVCC").unwrap_or_else(|_| "nvcc".into()))
        .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17"])
        .args(["-Xcompiler", "-fPIC", "-shared", "-o"])
        .arg(&lib)
        .arg(&api)
        .arg(&host_copy)
        .status()
        .expect("nvcc not found");
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=plonkish_cuda");
}
