"""Builds the in-tree native artefacts with nvcc for sm_100a (cross-compiles without a GPU):
  ppmmodelpangenome_b200/libppm_b200.so   the C-ABI engine (include/ppm_b200.h)
  ppmmodelpangenome_b200/inference        the reference-compatible CLI
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, os.environ.get("PPM_LIB_NAME", "libppm_b200.so"))
CLI = os.path.join(HERE, "inference")
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v"] + os.environ.get("PPM_EXTRA_NVCC_FLAGS", "").split()
HOSTCXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else None


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def lib_sources():
    return [os.path.join(CSRC, f) for f in ("host_model.cpp", "band_program.cpp", "simulator.cpp", "ppm_kernels.cu", "band_kernel.cu", "band_stream.cu", "pattern_kernels.cu", "band_sparse.cu", "ppm_engine.cu")]


def all_inputs():
    out = []
    for d in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        out += [os.path.join(d, f) for f in os.listdir(d)]
    return out


def build_lib(force=False, verbose=False):
    if not force and not _newer(LIB, all_inputs()):
        return LIB
    cmd = [NVCC] + ARCH + COMMON + (["-ccbin", HOSTCXX] if HOSTCXX else []) + ["-shared", "-o", LIB] + lib_sources()
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode:
        print(r.stdout + r.stderr)
    if r.returncode:
        raise RuntimeError("nvcc failed building libppm_b200.so")
    with open(os.path.join(HERE, "build_ptxas.log"), "w") as f:
        f.write(r.stdout + r.stderr)
    return LIB


def build_cli(force=False, verbose=False):
    src = os.path.join(CSRC, "ppm_cli.cpp")
    if not os.path.exists(src):
        return None
    build_lib(force=force, verbose=verbose)
    if not force and not _newer(CLI, all_inputs() + [LIB]):
        return CLI
    cmd = [HOSTCXX or "g++", "-O2", "-std=c++17", "-o", CLI, src, "-L" + HERE, "-lppm_b200", "-Wl,-rpath,$ORIGIN"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode:
        print(r.stdout + r.stderr)
    if r.returncode:
        raise RuntimeError("g++ failed building the inference CLI")
    return CLI


if __name__ == "__main__":
    build_lib(force=True, verbose=True)
    build_cli(force=True, verbose=True)
