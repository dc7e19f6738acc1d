"""Build libkcvae.so (nvcc, sm_100a) in-tree.  ``python trustedai-cl-vae-ad_b200/build.py``.

``build_emu()`` builds the g++ functional simulation of the same sources used by the
CPU-only kernel-logic tests (tests/emu/); the product never loads that library."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libkcvae.so")
EMU_DIR = os.path.join(ROOT, "tests", "emu")
EMU_LIB = os.path.join(EMU_DIR, "_build", "libkcvae_emu.so")

SOURCES = ["conv.cu", "dense.cu", "loss.cu", "model.cu", "frontend.cu"]
CUDA_ONLY_SOURCES = ["tc_conv.cu", "tc_gen.cu"]  # tcgen05 / TMA kernels: not emulated

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--use_fast_math=false", "-Xptxas", "-v"]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def _newer(target: str, deps) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def _all_deps():
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(ROOT, "include", "kcvae.h"))
    deps.append(os.path.abspath(__file__))
    return deps


def _run(cmd, log):
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log.append("$ " + " ".join(cmd) + "\n" + p.stdout)
    if p.returncode != 0:
        raise RuntimeError("build failed:\n" + "\n".join(log[-2:]))


def build(force: bool = False, verbose: bool = False) -> str:
    srcs = [s for s in SOURCES + CUDA_ONLY_SOURCES if os.path.exists(os.path.join(CSRC, s))]
    if not force and _newer(LIB, _all_deps()):
        return LIB
    nvcc = _nvcc()
    objdir = os.path.join(PKG, "build")
    os.makedirs(objdir, exist_ok=True)
    log = []
    flags = [f for f in NVCC_FLAGS if f != "--use_fast_math=false"]
    flags += os.environ.get("KCVAE_NVCC_EXTRA", "").split()   # development builds, e.g. -DKCVAE_TAIL_TIMING

    def one(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        _run([nvcc, *flags, "-I", CSRC, "-I", "/usr/include", "-c", os.path.join(CSRC, src), "-o", obj], log)
        return obj

    with ThreadPoolExecutor(max_workers=len(srcs)) as ex:
        objs = list(ex.map(one, srcs))
    _run([nvcc, "-shared", "-o", LIB, *objs, "-ldl"], log)  # cudart is linked statically (nvcc default)
    with open(os.path.join(objdir, "build.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    return LIB


def build_emu(force: bool = False, sanitize: bool = False) -> str:
    out = EMU_LIB if not sanitize else EMU_LIB.replace(".so", "_asan.so")
    if not force and _newer(out, _all_deps() + [os.path.join(EMU_DIR, "cuda_emu.h")]):
        return out
    os.makedirs(os.path.dirname(out), exist_ok=True)
    cmd = ["g++", "-std=c++17", "-O2", "-g", "-fPIC", "-shared", "-DKCVAE_EMU", "-I", EMU_DIR, "-I", CSRC,
           "-Wno-unused-parameter", "-Wno-attributes"]
    if sanitize:
        cmd += ["-fsanitize=address,undefined", "-fno-omit-frame-pointer"]
    for s in SOURCES:
        cmd += ["-x", "c++", os.path.join(CSRC, s)]
    cmd += ["-o", out]
    log = []
    _run(cmd, log)
    return out


if __name__ == "__main__":
    if "--emu" in sys.argv:
        print(build_emu(force=True))
    else:
        print(build(force=True, verbose="-v" in sys.argv))
