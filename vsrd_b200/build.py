"""Build libvsrd_b200.so (sm_100a) in-tree with nvcc.

    python -m vsrd_b200.build [--force] [--verbose]

Seven translation units are compiled in parallel (the field kernels are fully unrolled and take a
few minutes of ptxas time each) and linked into `vsrd_b200/libvsrd_b200.so`.  The .so is git-ignored
but travels to the GPU box with the source snapshot.
"""
from __future__ import annotations

import concurrent.futures
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ_DIR = os.path.join(HERE, "_obj")
LIB_PATH = os.path.join(HERE, "libvsrd_b200.so")
SOURCES = ["vsrd_render.cu", "vsrd_field_fwd.cu", "vsrd_field_umma.cu", "vsrd_field_bwd_umma.cu", "vsrd_field_bwd.cu", "vsrd_field_bwd_mma.cu", "vsrd_frame.cu", "vsrd_surface.cu", "vsrd_model.cu"]
HEADERS = ["vsrd_common.cuh", "vsrd_math.cuh", "vsrd_frag.cuh", "vsrd_umma.cuh", "vsrd_umma_field.cuh", "vsrd_frame_math.cuh", os.path.join("..", "..", "include", "vsrd_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; vsrd_b200 needs the CUDA 12.9 toolkit to build its kernels")
    return nvcc


def _newest_header_mtime() -> float:
    return max(os.path.getmtime(os.path.join(CSRC, h)) for h in HEADERS)


def _compile(src: str, force: bool, verbose: bool) -> str:
    obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
    path = os.path.join(CSRC, src)
    if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(path), _newest_header_mtime()):
        return obj
    cmd = [_nvcc(), *NVCC_FLAGS, "-c", path, "-o", obj]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    log = proc.stdout + proc.stderr
    with open(obj + ".log", "w") as f:
        f.write(log)
    if proc.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{log[-4000:]}")
    if verbose:
        print(log)
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ_DIR, exist_ok=True)
    with concurrent.futures.ThreadPoolExecutor(max_workers=len(SOURCES)) as pool:
        objs = list(pool.map(lambda s: _compile(s, force, verbose), SOURCES))
    if force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(o) > os.path.getmtime(LIB_PATH) for o in objs):
        cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "--shared", "-o", LIB_PATH, *objs]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        if proc.returncode != 0:
            raise RuntimeError(f"link failed:\n{proc.stdout}{proc.stderr}")
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
