// build.rs — compiles the CUDA library with nvcc for sm_100a and links it (north_star: "a thin
// extern "C" FFI crate built by build.rs with nvcc").  UNVERIFIED here: no cargo/rustc in the image;
// the same nvcc command is what __graft_entry__.build() runs.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..");
    let csrc = root.join("rust_pathtracer_b200/csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".into());
    let lib = out.join("libptb200.so");
    let status = Command::new(&nvcc)
        .args(["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
               "-prec-div=false", "-prec-sqrt=false", "-ftz=true", "-shared", "-Xcompiler", "-fPIC", "-o"])
        .arg(&lib)
        .arg(csrc.join("ptb_api.cu"))
        .status()
        .expect("nvcc not found: this crate has no CPU fallback");
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=ptb200");
    println!("cargo:rustc-link-lib=dylib=cudart");
    for f in ["ptb_api.cu", "ptb_device.cuh", "ptb_kernels.cuh", "ptb_wavefront.cuh", "ptb_wavefront_async.cuh", "ptb_stream.cuh"] {
        println!("cargo:rerun-if-changed={}", csrc.join(f).display());
    }
    println!("cargo:rerun-if-changed={}", root.join("include/ptb200.h").display());
}
