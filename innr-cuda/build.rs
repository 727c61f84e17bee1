// Builds libinnr_cuda.so for sm_100a from the CUDA sources of this repository (innr_b200/csrc/*.cu, header in include/)
// with the same flags as innr_b200/csrc/Makefile, or links a prebuilt one (feature `prebuilt` + $INNR_CUDA_LIB_DIR).
// Not compiled in the authoring environment (no cargo/rustc there).
use std::{env, fs, path::PathBuf, process::Command};

fn main() {
    let manifest = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap());
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    println!("cargo:rerun-if-env-changed=INNR_CUDA_LIB_DIR");
    if env::var("CARGO_FEATURE_PREBUILT").is_ok() {
        let dir = env::var("INNR_CUDA_LIB_DIR").expect("feature `prebuilt` needs INNR_CUDA_LIB_DIR");
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-lib=dylib=innr_cuda");
        return;
    }
    // the kernels live next to this crate: <repo>/innr_b200/csrc, header in <repo>/include
    let csrc = manifest.join("..").join("innr_b200").join("csrc");
    let mut sources: Vec<PathBuf> = fs::read_dir(&csrc)
        .unwrap_or_else(|e| panic!("cannot list {}: {e}", csrc.display()))
        .filter_map(|entry| entry.ok().map(|e| e.path()))
        .filter(|p| p.extension().map_or(false, |x| x == "cu"))
        .collect();
    sources.sort();
    assert!(!sources.is_empty(), "no .cu sources in {}", csrc.display());
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".into());
    let mut objects = Vec::new();
    for src in &sources {
        let obj = out.join(src.file_stem().unwrap()).with_extension("o");
        let status = Command::new(&nvcc)
            .args(["-O3", "-std=c++17", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a"])
            .args(["-Xcompiler", "-fPIC,-ffp-contract=off", "-c"])
            .arg(src)
            .arg("-o")
            .arg(&obj)
            .status()
            .expect("failed to run nvcc");
        assert!(status.success(), "nvcc failed on {}", src.display());
        println!("cargo:rerun-if-changed={}", src.display());
        objects.push(obj);
    }
    for header in ["common.cuh", "kernels.cuh", "tc_common.cuh"] {
        println!("cargo:rerun-if-changed={}", csrc.join(header).display());
    }
    println!("cargo:rerun-if-changed={}", manifest.join("..").join("include").join("innr_cuda.h").display());
    let lib = out.join("libinnr_cuda.so");
    let status = Command::new(&nvcc)
        .args(["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o"])
        .arg(&lib)
        .args(&objects)
        .status()
        .expect("failed to run nvcc");
    assert!(status.success(), "nvcc failed to link libinnr_cuda.so");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=innr_cuda");
}
