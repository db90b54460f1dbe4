"""Builds the C-ABI shared library (CUDA kernels for sm_100a) and the host tools, in-tree.

    python -m pixel_art_remaster_gpu_b200.build            # library + CLI
nvcc cross-compiles without a GPU.  The library is pixel_art_remaster_gpu_b200/libpixelart_b200.so.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
HOST = os.path.join(HERE, "host")
LIB = os.path.join(HERE, "libpixelart_b200.so")
CLI = os.path.join(HERE, "remaster_cli")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-O2",
              "-diag-suppress", "177"]


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def sources():
    cu = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu") or f.endswith(".cpp"))
    return [os.path.join(CSRC, f) for f in cu]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force=False, verbose=False):
    srcs = sources()
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    deps.append(os.path.join(ROOT, "include", "pixelart_b200.h"))
    if not force and not _stale(LIB, deps):
        return LIB
    objs = []
    build_dir = os.path.join(HERE, "build")
    os.makedirs(build_dir, exist_ok=True)
    procs = []
    for s in srcs:
        o = os.path.join(build_dir, os.path.basename(s) + ".o")
        objs.append(o)
        cmd = [_nvcc()] + ARCH + NVCC_FLAGS + \
              ["-I", os.path.join(ROOT, "include"), "-x", "cu", "-c", s, "-o", o]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (s, out))
        if verbose and out:
            print(out)
    cmd = [_nvcc()] + ARCH + ["-shared", "-o", LIB] + objs
    subprocess.run(cmd, check=True)
    return LIB


def build_cli(force=False):
    srcs = [os.path.join(HOST, f) for f in sorted(os.listdir(HOST)) if f.endswith(".cpp")] if os.path.isdir(HOST) else []
    if not srcs:
        return None
    deps = srcs + [os.path.join(HOST, f) for f in os.listdir(HOST) if f.endswith(".h")] + [LIB]
    if not force and not _stale(CLI, deps):
        return CLI
    cmd = ["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "include"), "-I", HOST] + srcs + \
          ["-o", CLI, "-L", HERE, "-lpixelart_b200", "-lz", "-Wl,-rpath,$ORIGIN"]
    subprocess.run(cmd, check=True)
    return CLI


def build_all(force=False, verbose=False):
    lib = build_library(force, verbose)
    build_cli(force)
    return lib


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv, verbose="-v" in sys.argv))
