"""In-tree builds: the sm_100a CUDA library, the synthetic-workload generator, the oracle and
(when /root/reference is present) the reference's own CUDA kernels into oracle/_ref/.

Everything is compiled with explicit nvcc / gcc command lines; outputs are git-ignored `.so`
files that travel to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "mach3_b200")
CSRC = os.path.join(PKG, "csrc")
SYNTH = os.path.join(PKG, "synth")
ORACLE = os.path.join(ROOT, "oracle")
REFERENCE = "/root/reference"

NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
GCC = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else (shutil.which("gcc") or "gcc")
GXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else (shutil.which("g++") or "g++")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _stale(out, srcs):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(s) > t for s in srcs if os.path.exists(s))


def _run(cmd, verbose):
    if verbose:
        print("+", " ".join(cmd), flush=True)
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError(f"build failed: {' '.join(cmd)}")
    if verbose and r.stdout.strip():
        print(r.stdout)
    return r.stdout


def build_synth(force=False, verbose=False):
    out = os.path.join(SYNTH, "libm3bsynth.so")
    srcs = [os.path.join(SYNTH, "m3b_synth.c"), os.path.join(SYNTH, "m3b_synth.h")]
    if force or _stale(out, srcs):
        _run([GCC, "-std=gnu11", "-O2", "-fopenmp", "-fPIC", "-shared", "-o", out, srcs[0], "-lm"], verbose)
    return out


def cuda_sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def build_cuda(force=False, verbose=False):
    """libm3b200.so: kernels + C-ABI (include/m3b200.h), hand-written for sm_100a."""
    exp = bool(os.environ.get("M3B_BUILD_EXPERIMENTS"))
    out = os.path.join(PKG, "libm3b200_exp.so" if exp else "libm3b200.so")
    srcs = cuda_sources()
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    deps.append(os.path.join(ROOT, "include", "m3b200.h"))
    if force or _stale(out, deps):
        cmd = [NVCC, *ARCH, "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC,-fvisibility=hidden",
               "-ccbin", GXX, "-I", os.path.join(ROOT, "include"), "-I", CSRC, "-shared", "-o", out, *srcs,
               "-Xptxas", "-v", "-lcudart"]
        if exp:      # A/B builds only (libm3b200_exp.so): legacy kernel variants + M3B_* environment knobs
            cmd.insert(1, "-DM3B_EXPERIMENTS")
        log = _run(cmd, verbose)
        os.makedirs(os.path.join(ROOT, "build"), exist_ok=True)          # git-ignored; registers / spills per kernel
        with open(os.path.join(ROOT, "build", "ptxas_exp.log" if exp else "ptxas.log"), "w") as f:
            f.write(log)
    return out


def build_oracle(force=False, verbose=False):
    out = os.path.join(ORACLE, "libm3oracle.so")
    if force or _stale(out, [os.path.join(ORACLE, "m3_oracle.c"), os.path.join(ORACLE, "Makefile")]):
        _run(["make", "-C", ORACLE, "-B" if force else "-s"], verbose)
    return out


def build_reference_gpu(force=False, verbose=False):
    """oracle/_ref/: the reference's OWN CUDA spline kernels (Splines/gpuSplineUtils.cu,
    Manager/gpuUtils.cu) compiled where they lie under /root/reference plus our harness.
    Only possible in the container that has /root/reference; the GPU box uses the prebuilt files."""
    mk = os.path.join(ORACLE, "ref_gpu", "Makefile")
    if not os.path.isdir(REFERENCE) or not os.path.exists(mk):
        return None
    _run(["make", "-C", os.path.dirname(mk), "-B" if force else "-s"], verbose)
    # the reference's own (header-only) binning code behind a C ABI: pins the oracle's FindBin / non-uniform binning
    hmk = os.path.join(ORACLE, "ref_host", "Makefile")
    if os.path.exists(hmk):
        _run(["make", "-j4", "-C", os.path.dirname(hmk), "-B" if force else "-s"], verbose)
    # the drop-in SMonolithGPU adapter, compiled against the reference's own header + the same harness
    amk = os.path.join(ROOT, "adapters", "Makefile")
    if os.path.exists(amk):
        _run(["make", "-C", os.path.dirname(amk), "-B" if force else "-s"], verbose)
    return os.path.join(ORACLE, "_ref")


def build_all(force=False, verbose=False):
    outs = [build_synth(force, verbose), build_cuda(force, verbose), build_oracle(force, verbose)]
    ref = build_reference_gpu(force, verbose)
    if ref:
        outs.append(ref)
    return outs


if __name__ == "__main__":
    for o in build_all(force="--force" in sys.argv, verbose=True):
        print("built", o)
