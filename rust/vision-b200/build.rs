// build.rs — compiles the CUDA C++ library (csrc/smh_b200.cu: kernels + extern "C" ABI) for
// sm_100a only and links it.  Replaces vision-gpu/build.rs:97-141 (nvcc --ptx for five
// compute capabilities + vendored PTX): no PTX, no JIT, no multi-arch dispatch, no NPP.
//
// UNCOMPILED here (no Rust toolchain in the build image).  SMH_B200_SRC points at this
// repository's checkout (the directory holding include/ and squad-mortar-helper_b200/csrc/).
use std::{env, path::PathBuf, process::Command};

fn main() {
    let src = PathBuf::from(env::var("SMH_B200_SRC").unwrap_or_else(|_| "../..".into()));
    let csrc = src.join("squad-mortar-helper_b200/csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    for f in ["smh_b200.cu", "kernels.cuh", "kernels_px.cuh", "kernels_lsd.cuh", "kernels_next.cuh", "kernels_icons.cuh", "smh_host_widen.cpp",
              "smh_pixel.h", "smh_synth.h", "smh_lsd_filter.h", "smh_px_plan.h"] {
        println!("cargo:rerun-if-changed={}", csrc.join(f).display());
    }
    println!("cargo:rerun-if-changed={}", src.join("include/smh_b200.h").display());
    println!("cargo:rerun-if-changed={}", src.join("include/smh_consts.h").display());

    let nvcc = which::which("nvcc").expect("nvcc (CUDA >= 12.8) is required: this backend has no CPU fallback");
    let obj = out.join("smh_b200.o");
    let status = Command::new(&nvcc)
        .args(["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17"])
        // results feed truncating casts: never contract a*b+c into an FMA
        .args(["-fmad=false", "-Xcompiler", "-fPIC,-ffp-contract=off"])
        .arg("-c").arg(csrc.join("smh_b200.cu")).arg("-o").arg(&obj)
        .status().expect("failed to run nvcc");
    assert!(status.success(), "nvcc failed");
    // host-only helper (AVX2 widening of the grayscale ui_map rows that cross the bus as L8)
    let obj2 = out.join("smh_host_widen.o");
    let status = Command::new(&nvcc)
        .args(["-O3", "-std=c++17", "-Xcompiler", "-fPIC"])
        .arg("-c").arg(csrc.join("smh_host_widen.cpp")).arg("-o").arg(&obj2)
        .status().expect("failed to run nvcc");
    assert!(status.success(), "nvcc (host file) failed");
    let lib = out.join("libsmh_b200_static.a");
    assert!(Command::new("ar").arg("crs").arg(&lib).arg(&obj).arg(&obj2).status().unwrap().success());

    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=static=smh_b200_static");
    let cuda_lib = nvcc.parent().unwrap().parent().unwrap().join("lib64");
    println!("cargo:rustc-link-search=native={}", cuda_lib.display());
    println!("cargo:rustc-link-lib=static=cudart_static");
    println!("cargo:rustc-link-lib=dylib=stdc++");
    println!("cargo:rustc-link-lib=dylib=dl");
    println!("cargo:rustc-link-lib=dylib=rt");
    println!("cargo:rustc-link-lib=dylib=pthread");
}
