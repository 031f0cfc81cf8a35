"""In-tree build of every native piece (explicit compiler invocations, no JIT cache).

  product   tiny_mp2v_dec_b200/_lib/libmp2v_b200.so      nvcc, sm_100a only (CUDA kernels + C ABI + host parser/decoder)
  tooling   tiny_mp2v_dec_b200/_lib/libmp2v_streamgen.so g++  (synthetic stream generator; tests + bench input)
  checker   oracle/libmp2v_oracle.so                      gcc  (CPU restatement, test infrastructure)
  checker   oracle/_ref/*                                 g++  (the unmodified reference, only where /root/reference exists)

Built objects are git-ignored but travel to the GPU box with the repository snapshot.
"""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "tiny_mp2v_dec_b200")
LIB = os.path.join(PKG, "_lib")
CSRC = os.path.join(PKG, "csrc")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

PRODUCT_LIB = os.path.join(LIB, "libmp2v_b200.so")
STREAMGEN_LIB = os.path.join(LIB, "libmp2v_streamgen.so")
ORACLE_LIB = os.path.join(ROOT, "oracle", "libmp2v_oracle.so")
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libmp2v_ref.so")

CUDA_SOURCES = ["recon_kernels.cu", "vlc_kernel.cu", "convert_kernel.cu", "recon_api.cu"]
HOST_SOURCES = ["host/mp2v_parser.cpp", "host/stream_index.cpp", "host/decoder.cpp", "host/decoder_capi.cpp", "host/numa.cpp"]


def _run(cmd, cwd=None):
    r = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("build step failed: %s\n%s" % (" ".join(cmd), r.stdout))
    return r.stdout


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources if os.path.exists(s))


def _deps(*dirs):
    out = []
    for d in dirs:
        for base, _, files in os.walk(d):
            out += [os.path.join(base, f) for f in files if f.endswith((".cu", ".cuh", ".cpp", ".h", ".c"))]
    return out


def build_product(force=False, verbose=False, defines=(), out=None):
    """nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo: cross-compiles without a GPU.
    `defines` / `out` build kernel-tuning variants next to the product (tools/dev only)."""
    os.makedirs(LIB, exist_ok=True)
    if out is not None:
        srcs = [os.path.join(CSRC, s) for s in CUDA_SOURCES + HOST_SOURCES]
        cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--shared",
               "-Xcompiler", "-fPIC,-fvisibility=hidden,-Wall,-pthread", "-I", os.path.join(ROOT, "include"), "-I", CSRC,
               "-I", os.path.join(CSRC, "host"), "-o", out] + ["-D" + d for d in defines] + srcs + ["-lpthread"]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        o = _run(cmd)
        if verbose:
            print(o)
        return out
    srcs = [os.path.join(CSRC, s) for s in CUDA_SOURCES + HOST_SOURCES if os.path.exists(os.path.join(CSRC, s))]
    deps = _deps(CSRC, os.path.join(ROOT, "include"))
    if not force and not _stale(PRODUCT_LIB, deps):
        return PRODUCT_LIB
    cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "--shared", "-Xcompiler", "-fPIC,-fvisibility=hidden,-Wall,-pthread",
           "-I", os.path.join(ROOT, "include"), "-I", CSRC, "-I", os.path.join(CSRC, "host"),
           "-o", PRODUCT_LIB] + srcs + ["-lpthread"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    out = _run(cmd)
    if verbose:
        print(out)
    return PRODUCT_LIB


def build_streamgen(force=False):
    os.makedirs(LIB, exist_ok=True)
    src = os.path.join(ROOT, "tools", "streamgen", "streamgen.cpp")
    deps = [src, os.path.join(ROOT, "tools", "streamgen", "streamgen.h")] + _deps(os.path.join(CSRC, "host"), os.path.join(ROOT, "include"))
    if force or _stale(STREAMGEN_LIB, deps):
        _run(["g++", "-std=c++17", "-O2", "-Wall", "-fPIC", "-shared", "-fvisibility=hidden", "-pthread",
              "-I", os.path.join(ROOT, "include"), "-I", os.path.join(CSRC, "host"), "-o", STREAMGEN_LIB, src])
    return STREAMGEN_LIB


def build_oracle(force=False):
    """The checker: C restatement always; the real reference only where its sources are mounted."""
    odir = os.path.join(ROOT, "oracle")
    if force:
        _run(["make", "-C", odir, "clean"])
    _run(["make", "-C", odir, "oracle"])
    if os.path.exists("/root/reference/src/core/decoder.cpp"):
        _run(["make", "-C", odir, "ref"])
    return ORACLE_LIB


def build_all(force=False, verbose=False):
    build_streamgen(force)
    build_product(force, verbose)
    build_oracle(force)      # after the product: oracle/_ref also holds the reference sample linked against it


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("built:", PRODUCT_LIB, STREAMGEN_LIB, ORACLE_LIB, REF_LIB if os.path.exists(REF_LIB) else "(no _ref)")
