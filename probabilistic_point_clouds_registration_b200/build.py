"""Builds the in-tree native artefacts with nvcc / g++ for sm_100a (cross-compiles without a GPU).

  csrc/libppcr_cuda.so        CUDA kernels + the C ABI of include/ppcr.h
  host/libppcr_registration.so the reference-facing C++ class (prob_point_cloud_registration::ProbPointCloudRegistration)
  host/prob_point_cloud_registration  the CLI

Run as `python -m probabilistic_point_clouds_registration_b200.build` or through __graft_entry__.build().
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
HOST = os.path.join(PKG, "host")
INCLUDE = os.path.join(ROOT, "include")

LIB_CUDA = os.path.join(CSRC, "libppcr_cuda.so")
LIB_REG = os.path.join(HOST, "libppcr_registration.so")
CLI_BIN = os.path.join(HOST, "prob_point_cloud_registration")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _gxx() -> str:
    for cand in ("/usr/bin/g++", shutil.which("g++")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("g++ not found")


def _stale(target: str, sources) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _sources(directory, exts):
    out = []
    for name in sorted(os.listdir(directory)):
        if name.endswith(exts):
            out.append(os.path.join(directory, name))
    return out


def build_cuda(force=False, verbose=False) -> str:
    srcs = _sources(CSRC, (".cu", ".cuh", ".h")) + [os.path.join(INCLUDE, "ppcr.h")]
    if force or _stale(LIB_CUDA, srcs):
        cmd = [_nvcc(), *NVCC_FLAGS, "-shared", "-ccbin", _gxx(), "-I", INCLUDE, "-o", LIB_CUDA,
               os.path.join(CSRC, "ppcr_capi.cu"), "-lcudart"]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        subprocess.check_call(cmd)
    return LIB_CUDA


def build_host(force=False) -> tuple[str, str]:
    reg_src = os.path.join(HOST, "prob_point_cloud_registration.cc")
    cli_src = os.path.join(HOST, "prob_point_cloud_registration_ex.cc")
    if not os.path.exists(reg_src):
        return "", ""
    hdrs = []
    for base, _, files in os.walk(INCLUDE):
        hdrs += [os.path.join(base, f) for f in files]
    hdrs += _sources(HOST, (".h", ".hpp"))
    common = [_gxx(), "-O2", "-std=c++17", "-fPIC", "-Wall", "-I", INCLUDE, "-I", os.path.join(INCLUDE, "ppcr_compat")]
    if force or _stale(LIB_REG, [reg_src, *hdrs]):
        subprocess.check_call([*common, "-shared", "-o", LIB_REG, reg_src, "-L", CSRC, "-lppcr_cuda",
                               "-Wl,-rpath,$ORIGIN/../csrc"])
    if os.path.exists(cli_src) and (force or _stale(CLI_BIN, [cli_src, reg_src, *hdrs])):
        subprocess.check_call([*common, "-o", CLI_BIN, cli_src, "-L", HOST, "-lppcr_registration", "-L", CSRC,
                               "-lppcr_cuda", "-Wl,-rpath,$ORIGIN:$ORIGIN/../csrc"])
    return LIB_REG, CLI_BIN


def build_all(force=False, verbose=False):
    build_cuda(force=force, verbose=verbose)
    build_host(force=force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB_CUDA)
