// build.rs for the upstream NRAPS crate: compiles the CUDA transport path with
// nvcc for sm_100a and links it.  (Shipped as source: this image has no Rust
// toolchain, so it is compile-checked only outside this container.)
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("NRAPS_B200_DIR").unwrap_or_else(|_| "../nraps_b200/csrc".into()));
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    // same lists as CU_SRCS / CPP_SRCS of nraps_b200/csrc/Makefile (mc_multi.cu, the NCCL driver, is optional)
    let cu = ["mc_source.cu", "mc_transport.cu", "mc_woodcock.cu", "mc_event.cu", "mc_finalize.cu", "mc_bank.cu", "mc_api.cu"];
    let cpp = ["host_input.cpp", "host_mesh.cpp", "host_output.cpp", "host_diffusion.cpp"];
    let mut objs = Vec::new();
    for f in cu.iter().chain(cpp.iter()) {
        let o = out.join(format!("{f}.o"));
        let st = Command::new("nvcc")
            .args(["-arch=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-fmad=false",
                   "-Xcompiler", "-fPIC,-ffp-contract=off", "-I"])
            .arg(root.join("../../include"))
            .arg("-c").arg(root.join(f)).arg("-o").arg(&o)
            .status().expect("nvcc not found");
        assert!(st.success(), "nvcc failed on {f}");
        objs.push(o);
        println!("cargo:rerun-if-changed={}", root.join(f).display());
    }
    let lib = out.join("libnraps_b200.a");
    assert!(Command::new("ar").arg("crs").arg(&lib).args(&objs).status().unwrap().success());
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=static=nraps_b200");
    println!("cargo:rustc-link-search=native=/usr/local/cuda/lib64");
    println!("cargo:rustc-link-lib=static=cudart_static");
    println!("cargo:rustc-link-lib=dylib=stdc++");
    println!("cargo:rustc-link-lib=dylib=dl");
    println!("cargo:rustc-link-lib=dylib=rt");
    println!("cargo:rustc-link-lib=dylib=pthread");
}
