"""ORACLE / REFERENCE BUILD — TEST INFRASTRUCTURE ONLY.

ctypes binding of oracle/_ref/libmpp_ref.so: the reference's OWN translation units for the hot
path (TPS_Astar.cpp, HolonomicBlend.cpp, DiffDriveCollisionGridBased.cpp, DiffDrive_C.cpp, the cost
evaluators, transform_pc_square_clipping.cpp, tp_obstacles_single_path.cpp, ...), compiled
unmodified from /root/reference against the MRPT stand-in oracle/ref/mrpt_shim.h and exported behind
the same `orc_*` C interface as the restatement (oracle/ref/ref_capi.cpp). The Python classes are the
ones of oracle/orc.py, bound to this library instead of liboracle.so:

    from oracle import ref
    R = ref.module()          # same API as `oracle.orc` (PTG, Planner, CostMap, eval_nodes, ...)

The library is built here when /root/reference exists (`make -C oracle ref`); on the GPU box the
prebuilt oracle/_ref/libmpp_ref.so travels with the snapshot and /root/reference is never read.
Only tests/, __graft_entry__ and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import ctypes as C
import importlib.util
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libmpp_ref.so")
REFERENCE_DIR = "/root/reference/mrpt_path_planning"
_mod = None


def build(force=False):
    """(Re)build libmpp_ref.so from the reference sources when they are present."""
    if os.path.isdir(REFERENCE_DIR):
        if force:
            subprocess.check_call(["make", "-C", _HERE, "-s", "clean"])
        subprocess.check_call(["make", "-C", _HERE, "-s", "ref"])
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("oracle/_ref/libmpp_ref.so is missing and %s is not available to build it"
                           % REFERENCE_DIR)
    return LIB_PATH


def available():
    try:
        build()
        return True
    except Exception:
        return False


def module():
    """A second instance of oracle/orc.py whose lib() is libmpp_ref.so."""
    global _mod
    if _mod is None:
        spec = importlib.util.spec_from_file_location("oracle._orc_on_reference", os.path.join(_HERE, "orc.py"))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        m._LIB_PATH = LIB_PATH
        m.build = build
        L = m.lib()
        assert L.orc_is_reference_build() == 1
        L.orc_ptgs_from_ini.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(C.c_void_p), C.c_int]
        m.KIND = "reference"
        _mod = m
    return _mod


def ptgs_from_ini(path, section="SelfDriving"):
    """TrajectoriesAndRobotShape::initFromConfigFile (the reference's loader) on an INI file."""
    m = module()
    arr = (C.c_void_p * 8)()
    n = m.lib().orc_ptgs_from_ini(path.encode(), section.encode(), arr, 8)
    if n < 0:
        raise RuntimeError("reference: " + m.lib().orc_last_error().decode())
    return [m.PTG(arr[i]) for i in range(n)]
