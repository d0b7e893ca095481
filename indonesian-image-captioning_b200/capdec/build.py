"""Build libcapdec.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python indonesian-image-captioning_b200/capdec/build.py [--force]

Objects go to csrc/build/, the shared library next to the `capdec` package
(indonesian-image-captioning_b200/libcapdec.so) so that it travels with the tree.
"""
import concurrent.futures
import hashlib
import os
import subprocess
import sys

PKG_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(PKG_ROOT, "csrc")
BUILD = os.path.join(CSRC, "build")
LIB = os.path.join(PKG_ROOT, "libcapdec.so")
SOURCES = ["capi.cu", "decoder.cu", "attention.cu", "pointwise.cu", "loss.cu", "gemm_simt.cu",
           "gemm_tc.cu", "beam.cu", "recur.cu", "optim.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]
NVCC_FLAGS += os.environ.get("CAPDEC_NVCC_FLAGS", "").split()      # e.g. -DCAPDEC_RECUR_FINE (debug stamps)


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _digest():
    h = hashlib.sha256()
    names = sorted(f for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh")))
    for n in names:
        with open(os.path.join(CSRC, n), "rb") as fh:
            h.update(n.encode())
            h.update(fh.read())
    with open(os.path.join(os.path.dirname(PKG_ROOT), "include", "capdec.h"), "rb") as fh:
        h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build_library(force=False, verbose=False):
    os.makedirs(BUILD, exist_ok=True)
    stamp = os.path.join(BUILD, "stamp")
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp):
        with open(stamp) as fh:
            if fh.read().strip() == dig:
                return LIB
    nvcc = _nvcc()

    def compile_one(src):
        obj = os.path.join(BUILD, src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        if verbose and r.stderr.strip():
            print(r.stderr, file=sys.stderr)
        return obj

    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                 "-Xcompiler", "-fPIC", "-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    with open(stamp, "w") as fh:
        fh.write(dig)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
