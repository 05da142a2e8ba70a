"""Boundary drift guard (CPU): the three descriptions of the C ABI must agree symbol by symbol --
include/slas_cuda.h (the contract), the `extern "C"` block of rust/src/backends/cuda.rs (what a slas maintainer
links against; never compiled here: no rustc), and the ctypes table of slas_b200/_lib.py -- and the nvcc step of
rust/build.rs must compile exactly the sources slas_b200/csrc/Makefile compiles.

`python tests/test_boundary_drift.py --emit` prints the extern block generated from the header (the text that
belongs in cuda.rs)."""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "slas_cuda.h")
RUST = os.path.join(ROOT, "rust", "src", "backends", "cuda.rs")
BUILD_RS = os.path.join(ROOT, "rust", "build.rs")
MAKEFILE = os.path.join(ROOT, "slas_b200", "csrc", "Makefile")

_SCALARS = {"int": "c_int", "long": "c_long", "size_t": "usize", "uint64_t": "u64", "float": "f32", "double": "f64",
            "void": "c_void", "char": "c_char"}


def _strip_comments(text: str) -> str:
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    return re.sub(r"//[^\n]*", " ", text)


def c_type_to_rust(ctype: str) -> str:
    """`const float *const *` -> `*const *const f32`, `void **` -> `*mut *mut c_void`, `size_t` -> `usize` ..."""
    toks = re.findall(r"[A-Za-z_][A-Za-z_0-9]*|\*", ctype)
    base = [t for t in toks if t not in ("const", "*")]
    assert len(base) == 1, ctype
    rust = _SCALARS[base[0]]
    # walk the declarator left to right: each '*' wraps what is to its left; a `const` directly before the '*'
    # (or leading, for the base type) makes that level *const
    const_pending = False
    seen_base = False
    for t in toks:
        if t == "const":
            const_pending = True
        elif t == "*":
            rust = ("*const " if const_pending else "*mut ") + rust
            const_pending = False
        else:
            seen_base = True
    assert seen_base
    return rust


def header_functions() -> dict:
    """name -> (rust return type, [rust argument types]) for every function include/slas_cuda.h declares"""
    text = _strip_comments(open(HEADER).read())
    out = {}
    for m in re.finditer(r"(?:^|\n)\s*((?:const\s+)?[A-Za-z_0-9]+\s*\*?)\s*(slas_cuda_[a-z_0-9]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        ret, name, args = m.group(1).strip(), m.group(2), " ".join(m.group(3).split())
        rargs = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                a = re.sub(r"\s*[A-Za-z_][A-Za-z_0-9]*$", "", a) if not a.endswith("*") else a  # drop the parameter name
                rargs.append(c_type_to_rust(a))
        out[name] = (c_type_to_rust(ret), rargs)
    return out


def rust_functions() -> dict:
    text = _strip_comments(open(RUST).read())
    block = re.search(r'extern\s+"C"\s*\{(.*?)\n\}', text, flags=re.S)
    assert block, 'no extern "C" block in cuda.rs'
    out = {}
    for m in re.finditer(r"fn\s+(slas_cuda_[a-z_0-9]+)\s*\(([^)]*)\)\s*(?:->\s*([^;]+))?;", block.group(1), flags=re.S):
        name, args, ret = m.group(1), " ".join(m.group(2).split()), (m.group(3) or "()").strip()
        rargs = []
        for a in filter(None, (x.strip() for x in args.split(","))):
            rargs.append(a.split(":", 1)[1].strip())
        norm = lambda t: t.replace("std::os::raw::", "").replace("std::ffi::", "")  # noqa: E731
        out[name] = (norm(ret), [norm(t) for t in rargs])
    return out


def emit_extern_block() -> str:
    lines = ['#[link(name = "slas_cuda")]', 'extern "C" {']
    text = _strip_comments(open(HEADER).read())
    names = {}
    for m in re.finditer(r"(slas_cuda_[a-z_0-9]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        pn = []
        args = " ".join(m.group(2).split())
        if args and args != "void":
            for i, a in enumerate(args.split(",")):
                mm = re.search(r"([A-Za-z_][A-Za-z_0-9]*)$", a.strip())
                pn.append(mm.group(1) if mm and not a.strip().endswith("*") else f"arg{i}")
        names[m.group(1)] = pn
    for name, (ret, args) in header_functions().items():
        params = ", ".join(f"{('r#' + p) if p in ('in', 'type', 'ref', 'box', 'move', 'fn') else p}: {t}" for p, t in zip(names[name], args))
        lines.append(f"    fn {name}({params}) -> {ret};")
    lines.append("}")
    return "\n".join(lines)


def test_every_header_symbol_is_bound_in_rust_with_the_same_signature():
    hdr, rs = header_functions(), rust_functions()
    assert len(hdr) > 90, f"header parse found only {len(hdr)} functions"
    missing = sorted(set(hdr) - set(rs))
    extra = sorted(set(rs) - set(hdr))
    assert not missing, f"cuda.rs does not bind: {missing}"
    assert not extra, f"cuda.rs binds symbols the header does not declare: {extra}"
    for name in hdr:
        assert hdr[name] == rs[name], f"{name}: header says {hdr[name]}, cuda.rs says {rs[name]}"


def test_ctypes_table_covers_the_header_with_the_same_arity():
    sys.path.insert(0, ROOT)
    from slas_b200 import _lib
    hdr = header_functions()
    assert set(hdr) == set(_lib.SIGNATURES), (sorted(set(hdr) - set(_lib.SIGNATURES)), sorted(set(_lib.SIGNATURES) - set(hdr)))
    for name, (_, args) in hdr.items():
        assert len(args) == len(_lib.SIGNATURES[name]), f"{name}: {len(args)} arguments in the header, {len(_lib.SIGNATURES[name])} in _lib.py"


def test_build_rs_compiles_what_the_makefile_compiles():
    mk = open(MAKEFILE).read()
    srcs = re.search(r"^SRCS\s*:=\s*(.*)$", mk, flags=re.M).group(1).split()
    rs = open(BUILD_RS).read()
    listed = re.search(r"KERNEL_SOURCES:\s*&\[&str\]\s*=\s*&\[(.*?)\];", rs, flags=re.S).group(1)
    rs_srcs = re.findall(r'"([a-z_0-9]+\.cu)"', listed)
    assert sorted(rs_srcs) == sorted(srcs), f"build.rs KERNEL_SOURCES != Makefile SRCS: {sorted(set(srcs) ^ set(rs_srcs))}"
    host = re.search(r"^HOSTSRCS\s*:=\s*(.*)$", mk, flags=re.M).group(1).split()
    rs_host = re.findall(r'"([a-z_0-9]+\.cpp)"', re.search(r"HOST_SOURCES:\s*&\[&str\]\s*=\s*&\[(.*?)\];", rs, flags=re.S).group(1))
    assert sorted(rs_host) == sorted(host), f"build.rs HOST_SOURCES != Makefile HOSTSRCS: {host} vs {rs_host}"
    # the pieces of the Makefile that are not plain `nvcc -c file.cu`: gemm_tiny.cu once per element type, and the
    # kernel text gemm_jit.cu embeds for NVRTC
    assert "gemm_tiny.cu" in rs and "SLAS_TINY_F64" in rs, "build.rs must compile gemm_tiny.cu twice (f32 and -DSLAS_TINY_F64)"
    assert "embed_jit_source.py" in rs and "gemm_small_jit_src.inc" in rs, "build.rs must run the embed step gemm_jit.cu needs"
    assert "staged_jit_src.inc" in rs and "staged_jit_src.inc" in mk, "build.rs must run the embed step staged_jit.cu needs"
    for inc in ("gemm_small_jit_src.inc", "staged_jit_src.inc"):   # ... over the same headers as the Makefile rule
        mk_hdrs = re.search(rf"embed_jit_source\.py \$@ (.*)$", mk[mk.index(inc + ":"):], flags=re.M).group(1).split()
        rs_hdrs = re.findall(r'"([a-z_0-9]+\.cuh)"', rs[rs.index(inc):][:400])
        assert mk_hdrs == rs_hdrs, (inc, mk_hdrs, rs_hdrs)
    for flag in ("arch=compute_100a,code=sm_100a", "-lineinfo", "-std=c++17"):
        assert flag in rs and flag in mk, flag
    # config knobs of the reference's build script style (build.rs:7-27 there): SLAS_* environment variables
    for knob in ("SLAS_CUDA_HOME", "SLAS_CUDA_ARCH", "SLAS_CUDA_DEVICES"):
        assert knob in rs, knob


if __name__ == "__main__":
    if "--emit" in sys.argv:
        print(emit_extern_block())
