"""In-tree build of libmcpb200.so (nvcc, sm_100a) — used by `__graft_entry__.build()` and lazily by
`capi.load_library()` when the shared object is missing or older than its sources.

The problem-specialised kernels are NOT built here: they are generated per problem and compiled by
NVRTC inside the library (`mcpb200_create`), with the cubins cached under `mcp_b200/_kcache/`.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmcpb200.so")
SOURCES = ["plan.cpp", "mcpb200.cpp", "static_kernels.cu"]
DEPS = SOURCES + ["plan.h", "kernel_template.cuh", os.path.join("..", "..", "include", "mcpb200.h")]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libmcpb200.so cannot be built (there is no CPU fallback)")


def _embed_template() -> str:
    """kernel_template.cuh → a C++ translation unit holding it as one string constant."""
    src = open(os.path.join(CSRC, "kernel_template.cuh"), encoding="utf-8").read()
    assert ')KTPL"' not in src
    out = os.path.join(CSRC, "kernel_template_embed.cpp")
    # split into chunks: some host compilers cap the length of one string literal
    chunks, step = [], 8000
    for i in range(0, len(src), step):
        chunks.append('R"KTPL(' + src[i:i + step] + ')KTPL"')
    text = ("// generated from kernel_template.cuh by mcp_b200/build.py — do not edit\n"
            "extern const char* mcpb200_kernel_template_source;\n"
            "const char* mcpb200_kernel_template_source =\n" + "\n".join(chunks) + ";\n")
    if not os.path.exists(out) or open(out, encoding="utf-8").read() != text:
        with open(out, "w", encoding="utf-8") as f:
            f.write(text)
    return out


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    embed = _embed_template()
    cmd = [_nvcc(), "-O2", "-std=c++17", *ARCH, "-lineinfo", "-Xcompiler", "-fPIC,-Wall", "-shared",
           "-o", LIB + ".tmp", *[os.path.join(CSRC, s) for s in SOURCES], embed,
           "-lnvrtc", "-ldl", "-lpthread", "-Xlinker", "-rpath=/usr/local/cuda/lib64"]
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libmcpb200.so:\n" + res.stdout + res.stderr)
    os.replace(LIB + ".tmp", LIB)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
