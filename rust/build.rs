// build.rs additions for `slas_backend::Cuda` -- UNVERIFIED SOURCE (no rustc/cargo in this image).
//
// The reference's build script (its build.rs:6-28) turns `SLAS_*` environment variables into consts of
// `$OUT_DIR/config.rs`.  That part stays as it is; a maintainer adds ONE call at the end of its `main`:
//
//     #[cfg(feature = "cuda")]
//     cuda_backend::compile_and_link(&out_dir, &mut f);     // `f` = the open config.rs file
//
// and this module.  New knobs, both optional: SLAS_CUDA_HOME (default /usr/local/cuda) and SLAS_CUDA_ARCH
// (default and only accepted value: sm_100a -- the kernels use tcgen05 / TMEM / bulk TMA).
#[cfg(feature = "cuda")]
mod cuda_backend {
    use std::env;
    use std::fs::File;
    use std::io::Write;
    use std::path::{Path, PathBuf};
    use std::process::Command;

    const KERNEL_SOURCES: &[&str] = &[
        "context.cu", "api.cu", "vector_ops.cu", "transpose.cu", "gemm_small.cu", "gemm_generic.cu",
        "gemm_large.cu", "gemm_ffma.cu", "gemm_tcgen05.cu", "comm.cu", "microbench.cu",
    ];

    pub fn compile_and_link(out_dir: &str, config_rs: &mut File) {
        for knob in ["SLAS_CUDA_HOME", "SLAS_CUDA_ARCH"] {
            println!("cargo:rerun-if-env-changed={knob}");
        }
        let cuda_home = env::var("SLAS_CUDA_HOME").unwrap_or_else(|_| "/usr/local/cuda".into());
        let arch = env::var("SLAS_CUDA_ARCH").unwrap_or_else(|_| "sm_100a".into());
        assert_eq!(arch, "sm_100a", "slas_backend::Cuda is written for sm_100a (B200) only");

        let csrc = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("cuda/csrc");
        let lib = Path::new(out_dir).join("libslas_cuda.so");
        let mut nvcc = Command::new(Path::new(&cuda_home).join("bin/nvcc"));
        nvcc.args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo"])
            .args(["-Xcompiler", "-fPIC", "-shared", "-o"])
            .arg(&lib);
        for src in KERNEL_SOURCES {
            let path = csrc.join(src);
            println!("cargo:rerun-if-changed={}", path.display());
            nvcc.arg(path);
        }
        nvcc.args(["-ldl", "-lpthread"]); // cudart is linked statically into the .so; NCCL is dlopen'ed at run time
        let status = nvcc.status().expect("nvcc not found (set SLAS_CUDA_HOME)");
        assert!(status.success(), "nvcc failed");

        writeln!(config_rs, "/// GPU architecture the Cuda backend was compiled for.\npub const CUDA_ARCH: &str = {arch:?};").unwrap();
        println!("cargo:rustc-link-search=native={out_dir}");
        println!("cargo:rustc-link-lib=dylib=slas_cuda");
    }
}
