// build.rs -- compiles the CUDA sources of r1cs-spartan_b200 for sm_100a and links them into the crate.
// A maintainer copies `rust/` next to the reference's Cargo.toml (or points SB_B200_SRC at r1cs-spartan_b200/csrc) and
// adds `build = "build.rs"` plus `mod ffi;`.  NOT compiled in the authoring environment: there is
// no cargo/rustc there; tests/test_wire_cpu.py::test_rust_shim_lists_every_export keeps this file and src/ffi.rs in
// step with include/spartan_b200.h and the Makefile's source list.
use std::process::Command;

const SOURCES: &[&str] = &["kernels_fr.cu", "msm.cu", "prover.cu", "comm_shm.cu", "indexer.cu"];

fn main() {
    let out = std::env::var("OUT_DIR").unwrap();
    let src_dir = std::env::var("SB_B200_SRC").unwrap_or_else(|_| "r1cs-spartan_b200/csrc".to_string());
    let mut objs = vec![];
    for s in SOURCES.iter() {
        let path = format!("{}/{}", src_dir, s);
        let o = format!("{}/{}.o", out, std::path::Path::new(s).file_stem().unwrap().to_str().unwrap());
        let st = Command::new("nvcc")
            .args(&["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
                    "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-c", &path, "-o", &o])
            .status().expect("nvcc not found");
        assert!(st.success(), "nvcc failed on {}", path);
        objs.push(o);
    }
    let lib = format!("{}/libspartan_b200.a", out);
    assert!(Command::new("ar").arg("crs").arg(&lib).args(&objs).status().unwrap().success());
    println!("cargo:rustc-link-search=native={}", out);
    println!("cargo:rustc-link-lib=static=spartan_b200");
    println!("cargo:rustc-link-search=native=/usr/local/cuda/lib64");
    println!("cargo:rustc-link-lib=cudart");
    println!("cargo:rustc-link-lib=stdc++");
    println!("cargo:rustc-link-lib=rt");          // shm_open / shm_unlink of the shared-memory exchange (comm_shm.cu)
    println!("cargo:rustc-link-lib=pthread");     // worker threads of the multi-GPU context
    println!("cargo:rerun-if-changed={}", src_dir);
}
