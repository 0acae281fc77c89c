"""In-tree builds of the native pieces (no JIT cache: the .so files travel with the repo snapshot).

  libgudni_b200.so  CUDA kernels + C-ABI shim        nvcc, sm_100a only
  libgudni_host.so  harness: wire-format producer    g++
"""
import os
import shutil
import subprocess

ROOT = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(ROOT)
CSRC = os.path.join(ROOT, "csrc")
LIB_CUDA = os.path.join(ROOT, "libgudni_b200.so")
LIB_HOST = os.path.join(ROOT, "libgudni_host.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    # parity contract with the oracle: IEEE f32, no FMA contraction, IEEE division / sqrt
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-pthread", "-shared",
]


def _host_cxx():
    # the image exports CXX=/opt/gcc/bin/g++, a wrapper that lacks libgomp.spec; prefer the system one
    return "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else (shutil.which("g++") or "g++")


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def _sources(dirpath, exts):
    out = []
    for base, _, files in os.walk(dirpath):
        out += [os.path.join(base, f) for f in files if f.endswith(exts)]
    return sorted(out)


def build_host(force=False, verbose=False):
    src = [os.path.join(CSRC, "host", "scene.cpp")]
    deps = _sources(os.path.join(CSRC, "host"), (".cpp", ".hpp")) + [os.path.join(REPO, "include", "gudni_b200.h")]
    if not force and _newer(LIB_HOST, deps):
        return LIB_HOST
    cmd = [_host_cxx(), "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-Wall", "-o", LIB_HOST] + src
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return LIB_HOST


def build_cuda(force=False, verbose=False, extra=()):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    src = _sources(CSRC, (".cu",))
    deps = src + _sources(CSRC, (".cuh", ".h")) + [os.path.join(REPO, "include", "gudni_b200.h")]
    if not force and _newer(LIB_CUDA, deps):
        return LIB_CUDA
    cmd = [nvcc] + NVCC_FLAGS + list(extra) + ["-I", os.path.join(REPO, "include"), "-o", LIB_CUDA] + src
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return LIB_CUDA


def build_oracle(force=False, verbose=False):
    odir = os.path.join(REPO, "oracle")
    cmd = ["make", "-C", odir] + (["-B"] if force else [])
    subprocess.run(cmd, check=True, stdout=None if verbose else subprocess.DEVNULL)
    return os.path.join(odir, "libgudni_oracle.so")


def build_reference(force=False, verbose=False):
    """oracle/_ref/libgudni_ref.so: the reference's kernel file compiled for the host (checker and CPU
    baseline only).  None when /root/reference is absent and nothing was prebuilt."""
    import sys
    if REPO not in sys.path:
        sys.path.insert(0, REPO)
    from oracle.refbuild import build_ref
    return build_ref.build(force=force, verbose=verbose)


LIB_STRAND_CHECK = os.path.join(REPO, "tests", "native", "libstrand_check.so")


def build_strand_check(force=False, verbose=False):
    """tests/native/libstrand_check.so: the level-3 kernels' per-shape logic (csrc/strand_build.cuh)
    compiled for the host — a test helper, see tests/native/strand_check.cpp."""
    src = os.path.join(REPO, "tests", "native", "strand_check.cpp")
    deps = [src, os.path.join(CSRC, "strand_build.cuh"), os.path.join(REPO, "include", "gudni_b200.h")]
    if not force and _newer(LIB_STRAND_CHECK, deps):
        return LIB_STRAND_CHECK
    cmd = [_host_cxx(), "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-Wall", "-o", LIB_STRAND_CHECK, src]
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return LIB_STRAND_CHECK


LIB_RASTER_EMU = os.path.join(REPO, "tests", "native", "libraster_emu.so")


def build_raster_emu(force=False, verbose=False, extra=(), suffix=""):
    """tests/native/libraster_emu.so: the raster kernels' own sources compiled with g++ against a host
    stand-in for the CUDA device language and run under a SIMT emulator — a test helper, see
    tests/native/raster_emu.cpp."""
    src = os.path.join(REPO, "tests", "native", "raster_emu.cpp")
    deps = [src, os.path.join(REPO, "tests", "native", "emu", "cuda_runtime.h"),
            os.path.join(REPO, "include", "gudni_b200.h")] + _sources(CSRC, (".cu", ".cuh"))
    # `extra` / `suffix`: a kernel variant (the same -D flags tools/variants.sh hands to nvcc), built beside the default
    out = LIB_RASTER_EMU.replace(".so", suffix + ".so") if suffix else LIB_RASTER_EMU
    if not force and _newer(out, deps):
        return out
    cmd = [_host_cxx(), "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-w"] + list(extra) + [
           "-I", os.path.join(REPO, "tests", "native", "emu"), "-I", os.path.join(REPO, "include"), "-o", out, src]
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return out
