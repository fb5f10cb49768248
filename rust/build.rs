// build.rs — UNTESTED SOURCE (no rustc/cargo in the build image).
//
// Compiles the CUDA sources of libpbrt_b200 for sm_100a with nvcc and links the result into the
// crate, as BASELINE.json's north star describes ("a thin extern \"C\" FFI layer built by build.rs
// with nvcc").  Drop this file next to the reference's Cargo.toml and add `build = "build.rs"`.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let root = PathBuf::from(env::var("PBRT_B200_ROOT").unwrap_or_else(|_| "../".into()));
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".into());
    let mut objs = Vec::new();
    for src in ["film.cu", "splat.cu", "splat_class.cu"] {
        let obj = out.join(src).with_extension("o");
        let status = Command::new(&nvcc)
            .args(["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo"])
            // Rust never contracts a*b+c; neither may the kernels (they ask for FMA explicitly)
            .args(["-fmad=false", "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-O2"])
            .arg("-c")
            .arg(root.join("pbrt_b200/csrc").join(src))
            .arg("-o")
            .arg(&obj)
            .status()
            .expect("nvcc not found: libpbrt_b200 has no CPU fallback");
        assert!(status.success(), "nvcc failed on {src}");
        objs.push(obj);
    }
    let lib = out.join("libpbrt_b200.a");
    let status = Command::new("ar").arg("crs").arg(&lib).args(&objs).status().unwrap();
    assert!(status.success());
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=static=pbrt_b200");
    println!("cargo:rustc-link-search=native=/usr/local/cuda/lib64");
    println!("cargo:rustc-link-lib=static=cudart_static");
    println!("cargo:rustc-link-lib=dylib=stdc++");
    println!("cargo:rustc-link-lib=dylib=dl");
    println!("cargo:rustc-link-lib=dylib=rt");
    println!("cargo:rerun-if-changed={}", root.join("pbrt_b200/csrc").display());
    println!("cargo:rerun-if-changed={}", root.join("include/pbrt_b200.h").display());
}
