"""Build libmpp_b200.so in-tree with nvcc for sm_100a (B200).

The shared library is the product; Python only loads it through ctypes. Run
``python -m mrpt_path_planning_b200.build`` or call ``build_library()``.
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libmpp_b200.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3,-Wall",
    "-ccbin", "/usr/bin/g++",
    "-shared", "-cudart", "static",
]


def sources():
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    srcs += sorted(glob.glob(os.path.join(CSRC, "host", "*.cpp")))
    return srcs


def _deps():
    d = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "host", "*.h"))
    d.append(os.path.join(HERE, "..", "include", "mppb.h"))
    d.append(os.path.abspath(__file__))
    return d


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(f) > t for f in _deps() if os.path.exists(f))


def build_library(force=False, verbose=False):
    """Compile every CUDA/C++ source of the package into lib/libmpp_b200.so."""
    if not force and not needs_build():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [NVCC] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + sources()
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("nvcc failed building libmpp_b200.so")
    if verbose:
        print(r.stdout)
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
