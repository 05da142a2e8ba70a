// build.rs additions for `slas_backend::Cuda` -- NON-BUILDING SOURCE (no rustc/cargo in this image; gated behind the
// `cuda` cargo feature).  tests/test_boundary_drift.py keeps KERNEL_SOURCES and the special steps below in lock-step
// with slas_b200/csrc/Makefile, which IS what builds and tests the library here.
//
// The reference's build script (its build.rs:6-28) turns `SLAS_*` environment variables into consts of
// `$OUT_DIR/config.rs`.  That part stays as it is; a maintainer adds ONE call at the end of its `main`:
//
//     #[cfg(feature = "cuda")]
//     cuda_backend::compile_and_link(&out_dir, &mut f);     // `f` = the open config.rs file
//
// and this module.  New knobs, all optional:
//   SLAS_CUDA_HOME     (default /usr/local/cuda)
//   SLAS_CUDA_ARCH     (default and only accepted value: sm_100a -- the kernels use tcgen05 / TMEM / bulk TMA)
//   SLAS_CUDA_DEVICES  (default 1) -> `config::CUDA_DEVICES`: how many GPUs `Cuda` drives from one process; the library
//                      also honours the environment variable of the same name at run time.
#[cfg(feature = "cuda")]
mod cuda_backend {
    use std::env;
    use std::fs::File;
    use std::io::Write;
    use std::path::{Path, PathBuf};
    use std::process::Command;

    // == SRCS of slas_b200/csrc/Makefile
    const KERNEL_SOURCES: &[&str] = &[
        "context.cu", "api.cu", "vector_ops.cu", "batched_cluster.cu", "reduce.cu", "fused_chain.cu", "transpose.cu", "gemm_small.cu",
        "gemm_jit.cu", "staged_jit.cu", "gemm_generic.cu", "gemm_large.cu", "gemm_large_f64.cu", "gemm_tcgen05.cu",
        "comm.cu", "multi.cu", "hostcopy.cu", "microbench.cu",
    ];
    // == HOSTSRCS of the Makefile
    const HOST_SOURCES: &[&str] = &["hostcopy_pool.cpp"];
    const NVCC_FLAGS: &[&str] = &[
        "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
    ];

    fn run(cmd: &mut Command, what: &str) {
        let status = cmd.status().unwrap_or_else(|e| panic!("{what}: {e} (set SLAS_CUDA_HOME)"));
        assert!(status.success(), "{what} failed");
    }

    pub fn compile_and_link(out_dir: &str, config_rs: &mut File) {
        for knob in ["SLAS_CUDA_HOME", "SLAS_CUDA_ARCH", "SLAS_CUDA_DEVICES"] {
            println!("cargo:rerun-if-env-changed={knob}");
        }
        let cuda_home = env::var("SLAS_CUDA_HOME").unwrap_or_else(|_| "/usr/local/cuda".into());
        let arch = env::var("SLAS_CUDA_ARCH").unwrap_or_else(|_| "sm_100a".into());
        assert_eq!(arch, "sm_100a", "slas_backend::Cuda is written for sm_100a (B200) only");
        let devices: usize = env::var("SLAS_CUDA_DEVICES").ok().and_then(|v| v.parse().ok()).unwrap_or(1);

        let csrc = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("cuda/csrc");
        let out = Path::new(out_dir);
        let nvcc = Path::new(&cuda_home).join("bin/nvcc");
        let mut objects: Vec<PathBuf> = Vec::new();

        // gemm_jit.cu embeds the kernel text it specialises at run time with NVRTC (Makefile: gemm_small_jit_src.inc)
        let inc = out.join("gemm_small_jit_src.inc");
        run(Command::new("python3").arg(csrc.join("embed_jit_source.py")).arg(&inc)
                .args(["async.cuh", "gemm_small_kernel.cuh", "gemm_tiny_kernel.cuh"].map(|h| csrc.join(h))),
            "embed_jit_source.py");
        // staged_jit.cu does the same for the staged small-object kernels (Makefile: staged_jit_src.inc)
        let inc = out.join("staged_jit_src.inc");
        run(Command::new("python3").arg(csrc.join("embed_jit_source.py")).arg(&inc)
                .args(["async.cuh", "divby.cuh", "staged_objects_kernel.cuh"].map(|h| csrc.join(h))),
            "embed_jit_source.py (staged objects)");

        for src in KERNEL_SOURCES {
            let path = csrc.join(src);
            println!("cargo:rerun-if-changed={}", path.display());
            let obj = out.join(src.replace(".cu", ".o"));
            run(Command::new(&nvcc).args(NVCC_FLAGS).arg("-I").arg(out).arg("-c").arg(&path).arg("-o").arg(&obj), src);
            objects.push(obj);
        }
        // plain host C++ (Makefile: HOSTSRCS): the copy-thread pool of the pageable-operand path
        for src in HOST_SOURCES {
            let obj = out.join(src.replace(".cpp", ".o"));
            run(Command::new("g++").args(["-O3", "-std=c++17", "-fPIC", "-c"]).arg(csrc.join(src)).arg("-o").arg(&obj), src);
            objects.push(obj);
        }
        // gemm_tiny.cu holds 64 kernels per element type: one object per type (Makefile: gemm_tiny_f32.o / gemm_tiny_f64.o)
        for (name, define) in [("gemm_tiny_f32.o", None), ("gemm_tiny_f64.o", Some("-DSLAS_TINY_F64"))] {
            let obj = out.join(name);
            let mut cmd = Command::new(&nvcc);
            cmd.args(NVCC_FLAGS);
            if let Some(d) = define { cmd.arg(d); }
            run(cmd.arg("-c").arg(csrc.join("gemm_tiny.cu")).arg("-o").arg(&obj), name);
            objects.push(obj);
        }
        let lib = out.join("libslas_cuda.so");
        // cudart is linked statically into the .so; NCCL, NVRTC and the driver entry points are bound at run time
        run(Command::new(&nvcc).args(["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o"]).arg(&lib)
                .args(&objects).args(["-ldl", "-lpthread"]),
            "link libslas_cuda.so");

        writeln!(config_rs, "/// GPU architecture the Cuda backend was compiled for.\npub const CUDA_ARCH: &str = {arch:?};").unwrap();
        writeln!(config_rs, "/// GPUs `slas_backend::Cuda` drives from one process (SLAS_CUDA_DEVICES).\npub const CUDA_DEVICES: usize = {devices};").unwrap();
        println!("cargo:rustc-link-search=native={out_dir}");
        println!("cargo:rustc-link-lib=dylib=slas_cuda");
    }
}
