"""Build libadafocus_b200.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension machinery).

The library is a plain C-ABI shared object (include/adafocus_b200.h); it links only the static CUDA runtime and
resolves cuTensorMapEncodeTiled through cudaGetDriverEntryPoint at run time, so it loads on a machine without a
GPU driver (symbol checks) and fails loudly at af_ctx_create there.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libadafocus_b200.so")
SOURCES = ["conv_gemm.cu", "gru_tc.cu", "stem_gemm.cu", "dwconv_tma.cu", "mbconv_fused.cu", "mbconv_rows.cu", "kernels.cu", "capi.cu"]
HEADERS = ["ptx.cuh", "gru_tc.cuh", "conv_gemm.cuh", "stem_gemm.cuh", "kernels.cuh", "dwconv_tma.cuh", "dw_strip.cuh", "mbconv_fused.cuh", "mbconv_rows.cuh", os.path.join("..", "..", "include", "adafocus_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-shared",
]


def _nvcc():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; cannot build libadafocus_b200.so")
    return nvcc


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [os.path.join(CSRC, h) for h in HEADERS]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    """Compile the CUDA sources into adafocus_b200/lib/libadafocus_b200.so. Returns the library path."""
    if not force and not needs_build():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH]
    cmd += [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        sys.stderr.write(res.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))


def source_hash():
    """sha256 over the CUDA sources + headers the library is built from: identifies a build in profiles/traffic.json."""
    import hashlib
    h = hashlib.sha256()
    for name in sorted(SOURCES) + sorted(HEADERS):
        path = os.path.join(CSRC, name)
        if os.path.exists(path):
            h.update(name.encode())
            with open(path, "rb") as f:
                h.update(f.read())
    return h.hexdigest()[:16]
