"""In-tree build of libpdr.so (hand-written sm_100a kernels + C ABI) with nvcc.

The shared object is written next to this file so it travels to the GPU box with the
repo snapshot; nothing is JIT-compiled at run time.
"""
import concurrent.futures
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
BUILD_DIR = os.path.join(ROOT, "build", "pdr")
LIB_PATH = os.path.join(HERE, "libpdr.so")

ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-I", CSRC,
                "-I", os.path.join(ROOT, "include")]

# Geometry kernels must reproduce the oracle's IEEE fp32 op order bit for bit, so the
# compiler may not contract a*b+c into an FMA there.
NO_FMA = {"geom_project.cu", "geom_splat.cu", "geom_unproject.cu", "geom_fill.cu", "geom_hpr.cu",
          "geom_attr.cu", "neighbors.cu"}


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; cannot build libpdr.so")
    return exe


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers_digest():
    h = hashlib.sha1()
    for d in (CSRC, os.path.join(ROOT, "include")):
        for f in sorted(os.listdir(d)):
            if f.endswith((".h", ".cuh")):
                with open(os.path.join(d, f), "rb") as fh:
                    h.update(fh.read())
    return h.hexdigest()


def _compile_one(src, hdr_digest, verbose):
    path = os.path.join(CSRC, src)
    with open(path, "rb") as fh:
        digest = hashlib.sha1(fh.read() + hdr_digest.encode()).hexdigest()
    obj = os.path.join(BUILD_DIR, src[:-3] + ".o")
    stamp = obj + ".sha1"
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == digest:
        return obj, False
    cmd = [_nvcc()] + ARCH_FLAGS + COMMON_FLAGS
    if src in NO_FMA:
        cmd += ["-fmad=false"]
    cmd += ["-c", path, "-o", obj]
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as fh:
        fh.write(digest)
    return obj, True


def build(force=False, verbose=False):
    """Compile every csrc/*.cu for sm_100a and link libpdr.so. Returns the library path."""
    os.makedirs(BUILD_DIR, exist_ok=True)
    if force:
        for f in os.listdir(BUILD_DIR):
            os.remove(os.path.join(BUILD_DIR, f))
    srcs = _sources()
    hdr = _headers_digest()
    with concurrent.futures.ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        results = list(ex.map(lambda s: _compile_one(s, hdr, verbose), srcs))
    objs = [o for o, _ in results]
    changed = any(c for _, c in results)
    if changed or not os.path.exists(LIB_PATH):
        cmd = [_nvcc()] + ARCH_FLAGS + ["-shared", "-o", LIB_PATH] + objs + ["-cudart", "static"]
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
