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
    "-Xcompiler", "-fPIC,-O3,-Wall,-ffp-contract=off",
    "-ccbin", "/usr/bin/g++",
]
LINK_FLAGS = ["-shared", "-cudart", "static", "-ccbin", "/usr/bin/g++",
              "-gencode", "arch=compute_100a,code=sm_100a", "-ldl"]
# Kernels whose arithmetic must round exactly like the reference's x86-64 build (no FMA contraction).
PER_FILE_FLAGS = {
    "tpobs_holo.cu": ["-fmad=false"],
    "edge_cost.cu": ["-fmad=false"],
    "expand.cu": ["-fmad=false"],
    "seams.cu": ["-fmad=false"],
}
OBJ_DIR = os.path.join(HERE, "lib", "obj")
BIN_DIR = os.path.join(HERE, "bin")
CLI_PATH = os.path.join(BIN_DIR, "path-planner-cli")
CLI_SRC = os.path.join(CSRC, "apps", "path_planner_cli.cpp")


def sources():
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    srcs += sorted(glob.glob(os.path.join(CSRC, "host", "*.cpp")))
    return srcs


def _deps():
    d = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "host", "*.h"))
    d.append(os.path.join(HERE, "..", "include", "mppb.h"))
    d.append(os.path.abspath(__file__))
    d.append(CLI_SRC)
    return d


def needs_build():
    if not os.path.exists(LIB_PATH) or not os.path.exists(CLI_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(f) > t for f in _deps() if os.path.exists(f))


def _compile_one(args):
    src, obj, verbose = args
    cmd = [NVCC] + NVCC_FLAGS + PER_FILE_FLAGS.get(os.path.basename(src), []) + (["-Xptxas", "-v"] if verbose else [])
    cmd += os.environ.get("MPPB_EXTRA_NVCC", "").split()  # experiments only (e.g. -DMPPB_GRID_MIN_BLOCKS=6)
    cmd += ["-c", "-o", obj, src]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    return src, r.returncode, r.stdout


def build_library(force=False, verbose=False):
    """Compile every CUDA/C++ source of the package (one object per source, in parallel) and link
    lib/libmpp_b200.so."""
    if not force and not needs_build():
        return LIB_PATH
    from concurrent.futures import ThreadPoolExecutor

    os.makedirs(OBJ_DIR, exist_ok=True)
    hdr_t = max(os.path.getmtime(f) for f in _deps() if os.path.exists(f) and not f.endswith((".cu", ".cpp")))
    jobs, objs = [], []
    for src in sources():
        obj = os.path.join(OBJ_DIR, os.path.basename(src) + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_t):
            jobs.append((src, obj, verbose))
    with ThreadPoolExecutor(max_workers=8) as ex:
        for src, rc, out in ex.map(_compile_one, jobs):
            if rc != 0:
                sys.stderr.write(out)
                raise RuntimeError("nvcc failed compiling %s" % src)
            if verbose:
                print(out)
    r = subprocess.run([NVCC] + LINK_FLAGS + ["-o", LIB_PATH] + objs, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("nvcc failed linking libmpp_b200.so")
    # the headless path-planner-cli front end, linked against the library next to it
    os.makedirs(BIN_DIR, exist_ok=True)
    r = subprocess.run(["/usr/bin/g++", "-std=c++17", "-O2", "-Wall", "-ffp-contract=off", "-o", CLI_PATH, CLI_SRC,
                        "-L" + LIB_DIR, "-lmpp_b200", "-Wl,-rpath,$ORIGIN/../lib", "-pthread", "-ldl"],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("g++ failed building path-planner-cli")
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
