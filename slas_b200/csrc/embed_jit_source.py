"""Build step: the device text NVRTC specialises at run time (gemm_jit.cu) = async.cuh + gemm_small_kernel.cuh +
gemm_tiny_kernel.cuh with their #include / #pragma once lines dropped and the few fixed-width typedefs NVRTC does not know prepended, written as
one C++ raw string literal.  Usage: embed_jit_source.py <out.inc> <header>..."""
import sys

PRELUDE = """// (generated from async.cuh, gemm_small_kernel.cuh and gemm_tiny_kernel.cuh by embed_jit_source.py)
typedef unsigned int uint32_t;
typedef unsigned long long uint64_t;
"""

out, headers = sys.argv[1], sys.argv[2:]
text = PRELUDE
for h in headers:
    for line in open(h):
        if line.lstrip().startswith("#include") or line.lstrip().startswith("#pragma once"):
            continue
        text += line
assert ')SLASJIT"' not in text
with open(out, "w") as f:
    f.write('R"SLASJIT(' + text + ')SLASJIT"\n')
