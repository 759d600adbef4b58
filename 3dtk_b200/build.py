"""Builds lib/libb200icp.so (the C-ABI library) for sm_100a with nvcc, in-tree.

Used by __graft_entry__.build(); nvcc cross-compiles without a GPU.  The library is the product:
the Python binding (3dtk_b200/__init__.py) refuses to work without it.
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "lib", "libb200icp.so")

SOURCES = ["b200icp.cu", "host_util.cpp", "lum_graph.cpp", "do_icp.cpp", "scan_files.cpp"]
HEADERS = ["common.cuh", "grid_build.cuh", "nn_search.cuh", "icp_kernels.cuh", "stream_kernels.cuh", "normals.cuh", "reduce.cuh", "solve.h",
           os.path.join("..", "..", "include", "b200icp.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3", "-shared",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libb200icp.so")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    for f in SOURCES + HEADERS:
        p = os.path.join(CSRC, f)
        if os.path.exists(p) and os.path.getmtime(p) > t:
            return True
    return False


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + \
          [os.path.join(CSRC, s) for s in SOURCES]
    env = dict(os.environ)
    # the image's default CXX (/opt/gcc) lacks some runtime specs; the distro g++ is the tested host compiler
    if os.path.exists("/usr/bin/g++"):
        cmd[1:1] = ["-ccbin", "/usr/bin/g++"]
    res = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
