// build.rs -- the reference's build script (build.rs:1-28) with the Cuda backend's additions.
// UNVERIFIED SOURCE (no rustc/cargo in this image).  New, all behind cargo feature `cuda`:
//   SLAS_CUDA_HOME (default /usr/local/cuda), SLAS_CUDA_ARCH (default sm_100a; the kernels use
//   tcgen05/TMEM/bulk-TMA and build for nothing else), and the nvcc compile + link step.
use std::env;
use std::fs::File;
use std::io::Write;
use std::path::{Path, PathBuf};
use std::process::Command;

fn main() {
    // ---- unchanged: env -> consts (reference build.rs:7-27) ----
    let slas_env_vars = [("BLAS_IN_DOT_IF_LEN_GE", "750")];
    let out_dir = env::var("OUT_DIR").unwrap();
    let dest_path = Path::new(&out_dir).join("config.rs");
    let mut f = File::create(&dest_path).unwrap();
    for (var, default_value) in slas_env_vars {
        println!("cargo:rerun-if-env-changed={var}");
        let value = env::var(&format!("SLAS_{var}")).unwrap_or(default_value.to_string());
        f.write_all(format!("
                /// value of environments variable `SLAS_{var}` during build.
                /// Default value is `{default_value:?}`.
                pub const {var}: usize = {value};
             ").as_bytes()).unwrap();
    }

    // ---- new: compile the CUDA kernels for sm_100a and link them ----
    if env::var("CARGO_FEATURE_CUDA").is_ok() {
        for var in ["SLAS_CUDA_HOME", "SLAS_CUDA_ARCH"] {
            println!("cargo:rerun-if-env-changed={var}");
        }
        let cuda_home = env::var("SLAS_CUDA_HOME").unwrap_or("/usr/local/cuda".to_string());
        let arch = env::var("SLAS_CUDA_ARCH").unwrap_or("sm_100a".to_string());
        assert_eq!(arch, "sm_100a", "slas_backend::Cuda is written for sm_100a (B200) only");
        let csrc = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("cuda/csrc");
        let sources = ["context.cu", "api.cu", "vector_ops.cu", "transpose.cu", "gemm_small.cu", "gemm_generic.cu",
                       "gemm_large.cu", "gemm_ffma.cu", "gemm_tcgen05.cu", "comm.cu"];
        let lib = Path::new(&out_dir).join("libslas_cuda.so");
        let mut cmd = Command::new(Path::new(&cuda_home).join("bin/nvcc"));
        cmd.args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
                  "-Xcompiler", "-fPIC", "-shared", "-o"]).arg(&lib);
        for s in sources {
            println!("cargo:rerun-if-changed={}", csrc.join(s).display());
            cmd.arg(csrc.join(s));
        }
        cmd.args(["-ldl", "-lpthread"]);
        assert!(cmd.status().expect("nvcc not found (set SLAS_CUDA_HOME)").success(), "nvcc failed");
        f.write_all(format!("
                /// GPU architecture the Cuda backend was compiled for.
                pub const CUDA_ARCH: &str = {arch:?};
             ").as_bytes()).unwrap();
        println!("cargo:rustc-link-search=native={out_dir}");
        println!("cargo:rustc-link-lib=dylib=slas_cuda");   // cudart is linked statically into the .so; NCCL is dlopen'ed
    }
}
