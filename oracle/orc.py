"""ORACLE — TEST INFRASTRUCTURE ONLY (ctypes binding of oracle/_build/liboracle.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module. The product (mrpt_path_planning_b200) never does.
Parity UNPINNED: the reference ships no golden vectors (DESIGN.md, section Oracle).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None

c_double_p = C.POINTER(C.c_double)
c_float_p = C.POINTER(C.c_float)


def build(force=False):
    """Compile the oracle with its Makefile (g++ -O3 -ffp-contract=off)."""
    if force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB_PATH)
        for f in os.listdir(_HERE)
        if f.endswith((".hpp", ".cpp", "Makefile"))
    ):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_last_error.restype = C.c_char_p
        L.orc_diffdrive_new.restype = C.c_void_p
        L.orc_diffdrive_new.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                                        C.c_double, c_double_p, c_double_p, C.c_int]
        L.orc_ptg_free.argtypes = [C.c_void_p]
        for name in ("orc_ptg_ref_distance", "orc_ptg_max_radius", "orc_ptg_step_duration"):
            getattr(L, name).restype = C.c_double
            getattr(L, name).argtypes = [C.c_void_p]
        L.orc_ptg_num_paths.argtypes = [C.c_void_p]
        L.orc_ptg_step_count.restype = C.c_long
        L.orc_ptg_step_count.argtypes = [C.c_void_p, C.c_int]
        L.orc_ptg_init_dist.restype = C.c_double
        L.orc_ptg_init_dist.argtypes = [C.c_void_p, C.c_int]
        L.orc_ptg_path_pose.argtypes = [C.c_void_p, C.c_int, C.c_uint, c_double_p]
        L.orc_ptg_inverse_map.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.POINTER(C.c_int),
                                          c_double_p]
        L.orc_ptg_step_for_dist.argtypes = [C.c_void_p, C.c_int, C.c_double, C.POINTER(C.c_uint)]
        L.orc_diffdrive_path.argtypes = [C.c_void_p, C.c_int, c_float_p]
        L.orc_diffdrive_grid_info.argtypes = [C.c_void_p, c_double_p, c_double_p, c_double_p, C.POINTER(C.c_int),
                                              C.POINTER(C.c_int), C.POINTER(C.c_long)]
        L.orc_diffdrive_grid_csr.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint16), c_float_p]
        L.orc_ptg_point_inside_shape.argtypes = [C.c_void_p, C.c_double, C.c_double]
        L.orc_local_obstacles.restype = C.c_long
        L.orc_local_obstacles.argtypes = [c_float_p, c_float_p, C.c_size_t, c_double_p, C.c_double, c_float_p,
                                          c_float_p]
        L.orc_eval_nodes.restype = C.c_double
        L.orc_eval_nodes.argtypes = [C.POINTER(C.c_void_p), C.c_int, c_float_p, c_float_p, C.c_size_t, c_double_p,
                                     C.c_size_t, C.c_double, c_double_p, C.c_int, C.c_int]
        _lib = L
    return _lib


def _err():
    return lib().orc_last_error().decode()


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a, t):
    return a.ctypes.data_as(t)


class PTG:
    """Handle on an oracle PTG object."""

    def __init__(self, handle):
        if not handle:
            raise RuntimeError("oracle: " + _err())
        self.h = C.c_void_p(handle)

    def __del__(self):
        try:
            if self.h:
                lib().orc_ptg_free(self.h)
                self.h = None
        except Exception:
            pass

    @property
    def num_paths(self):
        return lib().orc_ptg_num_paths(self.h)

    @property
    def ref_distance(self):
        return lib().orc_ptg_ref_distance(self.h)

    @property
    def max_radius(self):
        return lib().orc_ptg_max_radius(self.h)

    @property
    def step_duration(self):
        return lib().orc_ptg_step_duration(self.h)

    def step_count(self, k):
        n = lib().orc_ptg_step_count(self.h, k)
        if n < 0:
            raise RuntimeError("oracle: " + _err())
        return n

    def init_dist(self, k):
        return lib().orc_ptg_init_dist(self.h, k)

    def path_pose(self, k, step):
        out = np.zeros(4)
        if lib().orc_ptg_path_pose(self.h, k, step, _p(out, c_double_p)) != 0:
            raise RuntimeError("oracle: " + _err())
        return out  # x, y, phi, dist

    def inverse_map(self, x, y, tol=0.1):
        k = C.c_int(0)
        d = C.c_double(0)
        r = lib().orc_ptg_inverse_map(self.h, x, y, tol, C.byref(k), C.byref(d))
        if r < 0:
            raise RuntimeError("oracle: " + _err())
        return bool(r), k.value, d.value

    def step_for_dist(self, k, dist):
        s = C.c_uint(0)
        r = lib().orc_ptg_step_for_dist(self.h, k, dist, C.byref(s))
        if r < 0:
            raise RuntimeError("oracle: " + _err())
        return bool(r), s.value

    def inside_shape(self, x, y):
        return bool(lib().orc_ptg_point_inside_shape(self.h, x, y))

    # --- DiffDrive only ---
    def path(self, k):
        n = self.step_count(k)
        out = np.zeros((n, 7), dtype=np.float32)
        if lib().orc_diffdrive_path(self.h, k, _p(out, c_float_p)) != 0:
            raise RuntimeError("oracle: " + _err())
        return out

    def grid(self):
        xmin, ymin, res = C.c_double(), C.c_double(), C.c_double()
        nx, ny, ne = C.c_int(), C.c_int(), C.c_long()
        if lib().orc_diffdrive_grid_info(self.h, C.byref(xmin), C.byref(ymin), C.byref(res), C.byref(nx),
                                         C.byref(ny), C.byref(ne)) != 0:
            raise RuntimeError("oracle: " + _err())
        offsets = np.zeros(nx.value * ny.value + 1, dtype=np.uint32)
        ks = np.zeros(ne.value, dtype=np.uint16)
        ds = np.zeros(ne.value, dtype=np.float32)
        lib().orc_diffdrive_grid_csr(self.h, _p(offsets, C.POINTER(C.c_uint32)), _p(ks, C.POINTER(C.c_uint16)),
                                     _p(ds, c_float_p))
        return dict(xmin=xmin.value, ymin=ymin.value, res=res.value, nx=nx.value, ny=ny.value, offsets=offsets,
                    ks=ks, ds=ds)


def diffdrive_c(num_paths, ref_distance, resolution, v_max, w_max_rad, K, shape_x, shape_y,
                turning_radius_ref=0.10):
    sx, sy = _f64(shape_x), _f64(shape_y)
    return PTG(lib().orc_diffdrive_new(num_paths, ref_distance, resolution, v_max, w_max_rad, K,
                                       turning_radius_ref, _p(sx, c_double_p), _p(sy, c_double_p), len(sx)))


def local_obstacles(xs, ys, pose, clip):
    xs, ys, pose = _f32(xs), _f32(ys), _f64(pose)
    ox = np.zeros(len(xs), dtype=np.float32)
    oy = np.zeros(len(xs), dtype=np.float32)
    n = lib().orc_local_obstacles(_p(xs, c_float_p), _p(ys, c_float_p), len(xs), _p(pose, c_double_p), clip,
                                  _p(ox, c_float_p), _p(oy, c_float_p))
    if n < 0:
        raise RuntimeError("oracle: " + _err())
    return ox[:n], oy[:n]


def eval_nodes(ptgs, xs, ys, poses, clip, mode=0, nthreads=1):
    """Free TP-distance per (node, ptg, k): returns (out[B][sumK] float64, seconds)."""
    xs, ys = _f32(xs), _f32(ys)
    poses = _f64(poses).reshape(-1, 3)
    sum_k = sum(p.num_paths for p in ptgs)
    out = np.zeros((len(poses), sum_k), dtype=np.float64)
    arr = (C.c_void_p * len(ptgs))(*[p.h for p in ptgs])
    t = lib().orc_eval_nodes(arr, len(ptgs), _p(xs, c_float_p), _p(ys, c_float_p), len(xs), _p(poses, c_double_p),
                             len(poses), clip, _p(out, c_double_p), mode, nthreads)
    if t < 0:
        raise RuntimeError("oracle: " + _err())
    return out, t


# Robot shape and PTG parameters of share/ptgs_ackermann_vehicle.ini (values typed in here as
# data; the polygon vertices are read as float by the reference, TrajectoriesAndRobotShape.cpp:25-27).
ACKERMANN_SHAPE_X = np.array([-0.10, 0.00, 0.3, 0.3, 0.0, -0.10], dtype=np.float32).astype(np.float64)
ACKERMANN_SHAPE_Y = np.array([0.15, 0.15, 0.1, -0.1, -0.15, -0.15], dtype=np.float32).astype(np.float64)


def ackermann_ptgs(num_paths=121, ref_distance=5.0, resolution=0.05):
    w = 60.0 * np.pi / 180.0  # mrpt::DEG2RAD(x) = x * M_PI / 180
    return [diffdrive_c(num_paths, ref_distance, resolution, 1.0, w, +1.0, ACKERMANN_SHAPE_X, ACKERMANN_SHAPE_Y),
            diffdrive_c(num_paths, ref_distance, resolution, 1.0, w, -1.0, ACKERMANN_SHAPE_X, ACKERMANN_SHAPE_Y)]
