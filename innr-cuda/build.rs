// UNVERIFIED SOURCE: this environment has no Rust toolchain. Mirrors INTEGRATION.md section 1.
use std::{env, path::PathBuf, process::Command};
fn main() {
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let srcs = ["api.cu", "scan_f32.cu", "layout.cu", "hamming.cu", "ternary.cu", "u8.cu", "maxsim.cu", "maxsim_tc.cu", "knn_tc.cu"];
    let mut objs = vec![];
    for s in srcs {
        let o = out.join(s).with_extension("o");
        let st = Command::new("nvcc")
            .args(["-O3", "-std=c++17", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
                   "-Xcompiler", "-fPIC,-ffp-contract=off", "-c"])
            .arg(format!("csrc/{s}")).arg("-o").arg(&o).status().expect("nvcc");
        assert!(st.success(), "nvcc failed on {s}");
        objs.push(o);
        println!("cargo:rerun-if-changed=csrc/{s}");
    }
    let so = out.join("libinnr_cuda.so");
    assert!(Command::new("nvcc").args(["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o"])
        .arg(&so).args(&objs).status().unwrap().success());
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=innr_cuda");
}
