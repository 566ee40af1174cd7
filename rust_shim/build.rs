// build.rs a sparkl maintainer would add to crates/sparkl3d (feature "cuda-b200"): compiles the hand-written sm_100a
// kernels + the C ABI with nvcc and links them, replacing crates-ptx/ (rustc_codegen_nvvm) and resources/*.ptx.
// NOT compiled in this repository's environment (no cargo/rustc in the image); kept as the reference-side binding.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let csrc = PathBuf::from(env::var("SPARKL_B200_CSRC").unwrap_or_else(|_| "../sparkl_b200/csrc".into()));
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    // one source tree, two crates — like the reference's `dim2` / `dim3` features (crates/sparkl2d, crates/sparkl3d): the
    // sparkl2d crate builds the same host sequencing / sort / ABI sources with -DSPK_DIM=2 and the dim2 stencil kernels
    let dim2 = env::var("CARGO_FEATURE_DIM2").is_ok();
    let sources: &[&str] = if dim2 {
        &["spk_api.cu", "spk_sort.cu", "spk_g2p2g2d.cu", "spk_grid2d.cu", "spk_dist.cu", "spk_fluids.cu", "spk_migrate.cu"]
    } else {
        &["spk_api.cu", "spk_sort.cu", "spk_g2p2g.cu", "spk_grid.cu", "spk_dist.cu", "spk_fluids.cu", "spk_migrate.cu"]
    };
    let mut objects = vec![];
    for s in sources {
        let obj = out.join(s.replace(".cu", ".o"));
        let status = Command::new("nvcc")
            .arg(if dim2 { "-DSPK_DIM=2" } else { "-DSPK_DIM=3" })
            .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-c"])
            .arg(csrc.join(s))
            .arg("-o")
            .arg(&obj)
            .status()
            .expect("nvcc not found");
        assert!(status.success(), "nvcc failed on {s}");
        objects.push(obj);
        println!("cargo:rerun-if-changed={}", csrc.join(s).display());
    }
    let lib = out.join("libsparkl_b200.a");
    assert!(Command::new("ar").arg("crs").arg(&lib).args(&objects).status().unwrap().success());
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=static=sparkl_b200");
    println!("cargo:rustc-link-search=native=/usr/local/cuda/lib64");
    println!("cargo:rustc-link-lib=cudart");
    println!("cargo:rustc-link-lib=stdc++");
}
