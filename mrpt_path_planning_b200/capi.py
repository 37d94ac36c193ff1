"""ctypes binding of libmpp_b200.so (the C ABI declared in include/mppb.h).

This module is plumbing for tests and benchmarks: the product is the shared library. It never
falls back to a CPU implementation — if the library or a CUDA device is missing, it raises.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

_lib = None

c_double_p = C.POINTER(C.c_double)
c_float_p = C.POINTER(C.c_float)
c_u32_p = C.POINTER(C.c_uint32)
c_u16_p = C.POINTER(C.c_uint16)


class GridDesc(C.Structure):
    _fields_ = [
        ("num_paths", C.c_int32),
        ("ref_distance", C.c_double),
        ("grid_x_min", C.c_double),
        ("grid_y_min", C.c_double),
        ("grid_res", C.c_double),
        ("grid_nx", C.c_int32),
        ("grid_ny", C.c_int32),
        ("cell_offsets", c_u32_p),
        ("entry_k", c_u16_p),
        ("entry_d", c_float_p),
        ("shape_nverts", C.c_int32),
        ("shape_x", c_double_p),
        ("shape_y", c_double_p),
        ("max_radius", C.c_double),
    ]


class DiffDriveCParams(C.Structure):
    _fields_ = [
        ("num_paths", C.c_int32),
        ("ref_distance", C.c_double),
        ("resolution", C.c_double),
        ("v_max_mps", C.c_double),
        ("w_max_rps", C.c_double),
        ("K", C.c_double),
        ("turning_radius_ref", C.c_double),
        ("shape_nverts", C.c_int32),
        ("shape_x", c_double_p),
        ("shape_y", c_double_p),
    ]


# name -> (restype, argtypes); every symbol include/mppb.h declares must be listed here
_SIGS = {
    "mppb_version": (C.c_char_p, []),
    "mppb_last_global_error": (C.c_char_p, []),
    "mppb_ctx_create": (C.c_int, [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]),
    "mppb_ctx_destroy": (None, [C.c_void_p]),
    "mppb_last_error": (C.c_char_p, [C.c_void_p]),
    "mppb_sync": (C.c_int, [C.c_void_p]),
    "mppb_stream": (C.c_void_p, [C.c_void_p]),
    "mppb_launch_count": (C.c_uint64, [C.c_void_p]),
    "mppb_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p)]),
    "mppb_host_free": (C.c_int, [C.c_void_p]),
    "mppb_timer_begin": (C.c_int, [C.c_void_p]),
    "mppb_timer_end": (C.c_int, [C.c_void_p, c_float_p]),
    "mppb_measure_fp32_tflops": (C.c_int, [C.c_void_p, c_double_p]),
    "mppb_measure_fp64_tflops": (C.c_int, [C.c_void_p, c_double_p]),
    "mppb_set_obstacles": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_double, C.c_int]),
    "mppb_set_obstacles_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_double, C.c_int]),
    "mppb_num_obstacles": (C.c_size_t, [C.c_void_p]),
    "mppb_clear_ptgs": (C.c_int, [C.c_void_p]),
    "mppb_set_ptg_grid": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(GridDesc)]),
    "mppb_num_ptgs": (C.c_int, [C.c_void_p]),
    "mppb_total_paths": (C.c_int, [C.c_void_p]),
    "mppb_eval_nodes": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_double, C.c_void_p]),
    "mppb_eval_nodes_device": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_double, C.c_void_p]),
    "mppb_nodes_xycs_from_poses": (None, [C.c_size_t, C.c_void_p, C.c_void_p]),
    "mppb_ptg_create_diffdrive_c": (C.c_int, [C.POINTER(DiffDriveCParams), C.POINTER(C.c_void_p)]),
    "mppb_ptg_destroy": (None, [C.c_void_p]),
    "mppb_ptg_last_error": (C.c_char_p, []),
    "mppb_ptg_num_paths": (C.c_int, [C.c_void_p]),
    "mppb_ptg_ref_distance": (C.c_double, [C.c_void_p]),
    "mppb_ptg_step_duration": (C.c_double, [C.c_void_p]),
    "mppb_ptg_max_radius": (C.c_double, [C.c_void_p]),
    "mppb_ptg_step_count": (C.c_int64, [C.c_void_p, C.c_int]),
    "mppb_ptg_path_pose": (C.c_int, [C.c_void_p, C.c_int, C.c_uint32, c_double_p]),
    "mppb_ptg_step_for_dist": (C.c_int, [C.c_void_p, C.c_int, C.c_double, c_u32_p]),
    "mppb_ptg_inverse_map": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_double, C.POINTER(C.c_int),
                                       c_double_p]),
    "mppb_ptg_init_dist": (C.c_double, [C.c_void_p, C.c_int]),
    "mppb_ptg_grid_export": (C.c_int, [C.c_void_p, C.POINTER(GridDesc)]),
    "mppb_ctx_add_ptg": (C.c_int, [C.c_void_p, C.c_void_p]),
}


def lib_path():
    return _build.LIB_PATH


def lib():
    """Load libmpp_b200.so (building it first if sources are newer). Raises if unavailable."""
    global _lib
    if _lib is None:
        if not os.path.exists(_build.LIB_PATH) or _build.needs_build():
            _build.build_library()
        L = C.CDLL(_build.LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)  # AttributeError if the library does not export it
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


class MppbError(RuntimeError):
    pass


def _ptr(a):
    """Device or host address of a numpy array / torch tensor / int."""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    if isinstance(a, np.ndarray):
        return C.c_void_p(a.ctypes.data)
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    raise TypeError("unsupported buffer type %r" % type(a))


class PTG:
    """Host-side PTG object (mirror of the reference's ptg_t)."""

    def __init__(self, handle):
        self.h = C.c_void_p(handle)

    @staticmethod
    def diffdrive_c(num_paths, ref_distance, resolution, v_max_mps, w_max_rps, K, shape_x, shape_y,
                    turning_radius_ref=0.10):
        sx = np.ascontiguousarray(shape_x, dtype=np.float64)
        sy = np.ascontiguousarray(shape_y, dtype=np.float64)
        p = DiffDriveCParams(num_paths, ref_distance, resolution, v_max_mps, w_max_rps, K, turning_radius_ref,
                             len(sx), sx.ctypes.data_as(c_double_p), sy.ctypes.data_as(c_double_p))
        h = C.c_void_p()
        if lib().mppb_ptg_create_diffdrive_c(C.byref(p), C.byref(h)) != 0:
            raise MppbError(lib().mppb_ptg_last_error().decode())
        return PTG(h.value)

    def __del__(self):
        try:
            if self.h:
                lib().mppb_ptg_destroy(self.h)
                self.h = None
        except Exception:
            pass

    @property
    def num_paths(self):
        return lib().mppb_ptg_num_paths(self.h)

    @property
    def ref_distance(self):
        return lib().mppb_ptg_ref_distance(self.h)

    @property
    def step_duration(self):
        return lib().mppb_ptg_step_duration(self.h)

    @property
    def max_radius(self):
        return lib().mppb_ptg_max_radius(self.h)

    def step_count(self, k):
        n = lib().mppb_ptg_step_count(self.h, k)
        if n < 0:
            raise MppbError(lib().mppb_ptg_last_error().decode())
        return n

    def path_pose(self, k, step):
        out = np.zeros(4)
        if lib().mppb_ptg_path_pose(self.h, k, step, out.ctypes.data_as(c_double_p)) != 0:
            raise MppbError(lib().mppb_ptg_last_error().decode())
        return out

    def step_for_dist(self, k, dist):
        s = C.c_uint32(0)
        r = lib().mppb_ptg_step_for_dist(self.h, k, dist, C.byref(s))
        if r < 0:
            raise MppbError(lib().mppb_ptg_last_error().decode())
        return bool(r), s.value

    def inverse_map(self, x, y, tol=0.1):
        k, d = C.c_int(0), C.c_double(0)
        r = lib().mppb_ptg_inverse_map(self.h, x, y, tol, C.byref(k), C.byref(d))
        if r < 0:
            raise MppbError(lib().mppb_ptg_last_error().decode())
        return bool(r), k.value, d.value

    def init_dist(self, k):
        return lib().mppb_ptg_init_dist(self.h, k)

    def init_dists(self):
        return np.array([self.init_dist(k) for k in range(self.num_paths)])

    def grid(self):
        d = GridDesc()
        if lib().mppb_ptg_grid_export(self.h, C.byref(d)) != 0:
            raise MppbError(lib().mppb_ptg_last_error().decode())
        nc = d.grid_nx * d.grid_ny
        offsets = np.ctypeslib.as_array(d.cell_offsets, shape=(nc + 1,)).copy()
        ne = int(offsets[-1])
        ks = np.ctypeslib.as_array(d.entry_k, shape=(ne,)).copy() if ne else np.zeros(0, np.uint16)
        ds = np.ctypeslib.as_array(d.entry_d, shape=(ne,)).copy() if ne else np.zeros(0, np.float32)
        return dict(xmin=d.grid_x_min, ymin=d.grid_y_min, res=d.grid_res, nx=d.grid_nx, ny=d.grid_ny,
                    offsets=offsets, ks=ks, ds=ds, max_radius=d.max_radius, num_paths=d.num_paths)


class Context:
    """One libmpp_b200 context = one CUDA device + one stream."""

    def __init__(self, device=0, stream=None):
        h = C.c_void_p()
        rc = lib().mppb_ctx_create(device, C.c_void_p(stream) if stream else None, C.byref(h))
        if rc != 0:
            raise MppbError(lib().mppb_last_global_error().decode())
        self.h = h
        self._keep = []

    def close(self):
        if self.h:
            lib().mppb_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise MppbError("%s (rc=%d)" % (lib().mppb_last_error(self.h).decode(), rc))

    def sync(self):
        self._ck(lib().mppb_sync(self.h))

    @property
    def stream(self):
        return lib().mppb_stream(self.h)

    @property
    def launch_count(self):
        return lib().mppb_launch_count(self.h)

    def timer_begin(self):
        self._ck(lib().mppb_timer_begin(self.h))

    def timer_end(self):
        ms = C.c_float(0)
        self._ck(lib().mppb_timer_end(self.h, C.byref(ms)))
        return ms.value

    def measure_fp32_tflops(self):
        v = C.c_double(0)
        self._ck(lib().mppb_measure_fp32_tflops(self.h, C.byref(v)))
        return v.value

    def measure_fp64_tflops(self):
        v = C.c_double(0)
        self._ck(lib().mppb_measure_fp64_tflops(self.h, C.byref(v)))
        return v.value

    # -- inputs --
    def set_obstacles(self, xs, ys, bin_size=0.0, append=False):
        xs = np.ascontiguousarray(xs, dtype=np.float32)
        ys = np.ascontiguousarray(ys, dtype=np.float32)
        assert xs.shape == ys.shape
        self._ck(lib().mppb_set_obstacles(self.h, _ptr(xs), _ptr(ys), xs.size, bin_size, int(append)))

    def set_obstacles_device(self, d_xs, d_ys, n, bin_size=0.0, append=False):
        self._ck(lib().mppb_set_obstacles_device(self.h, _ptr(d_xs), _ptr(d_ys), n, bin_size, int(append)))

    @property
    def num_obstacles(self):
        return lib().mppb_num_obstacles(self.h)

    def clear_ptgs(self):
        self._ck(lib().mppb_clear_ptgs(self.h))

    def add_ptg(self, ptg):
        if lib().mppb_ctx_add_ptg(self.h, ptg.h) != 0:
            msg = lib().mppb_last_error(self.h).decode() or lib().mppb_ptg_last_error().decode()
            raise MppbError(msg)

    def set_ptg_grid(self, g, shape_x, shape_y, ref_distance):
        """Upload raw collision-grid tables (dict as returned by PTG.grid())."""
        off = np.ascontiguousarray(g["offsets"], dtype=np.uint32)
        ks = np.ascontiguousarray(g["ks"], dtype=np.uint16)
        ds = np.ascontiguousarray(g["ds"], dtype=np.float32)
        sx = np.ascontiguousarray(shape_x, dtype=np.float64)
        sy = np.ascontiguousarray(shape_y, dtype=np.float64)
        d = GridDesc(g["num_paths"], ref_distance, g["xmin"], g["ymin"], g["res"], g["nx"], g["ny"],
                     off.ctypes.data_as(c_u32_p), ks.ctypes.data_as(c_u16_p), ds.ctypes.data_as(c_float_p),
                     len(sx), sx.ctypes.data_as(c_double_p), sy.ctypes.data_as(c_double_p), g["max_radius"])
        self._ck(lib().mppb_set_ptg_grid(self.h, self.num_ptgs, C.byref(d)))

    @property
    def num_ptgs(self):
        return lib().mppb_num_ptgs(self.h)

    @property
    def total_paths(self):
        return lib().mppb_total_paths(self.h)

    # -- stage (a)+(b) --
    def eval_nodes(self, poses, clip_dist, out=None):
        """Host entry point: poses [n][3] float64 -> out [n][total_paths] float32 (+inf = unconstrained)."""
        poses = np.ascontiguousarray(poses, dtype=np.float64).reshape(-1, 3)
        n = poses.shape[0]
        if out is None:
            out = np.empty((n, self.total_paths), dtype=np.float32)
        self._ck(lib().mppb_eval_nodes(self.h, n, _ptr(poses), clip_dist, _ptr(out)))
        return out

    def eval_nodes_device(self, n, d_nodes_xycs, clip_dist, d_out):
        self._ck(lib().mppb_eval_nodes_device(self.h, n, _ptr(d_nodes_xycs), clip_dist, _ptr(d_out)))


def nodes_xycs_from_poses(poses):
    poses = np.ascontiguousarray(poses, dtype=np.float64).reshape(-1, 3)
    out = np.empty((poses.shape[0], 4), dtype=np.float64)
    lib().mppb_nodes_xycs_from_poses(poses.shape[0], _ptr(poses), _ptr(out))
    return out


def free_distance(out_row_f32, init_dists):
    """freeDistance = min(init_k, (double)out) as tp_obstacles_single_path returns it."""
    return np.minimum(np.asarray(init_dists, dtype=np.float64), out_row_f32.astype(np.float64))


def pinned_empty(shape, dtype):
    """numpy array over pinned host memory from mppb_host_alloc (freed when the array is collected)."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape))
    p = C.c_void_p()
    if lib().mppb_host_alloc(n * dtype.itemsize, C.byref(p)) != 0:
        raise MppbError(lib().mppb_last_global_error().decode())
    buf = (C.c_byte * (n * dtype.itemsize)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype).reshape(shape)
    _PINNED[arr.ctypes.data] = p
    return arr


_PINNED = {}


# PTG set of share/ptgs_ackermann_vehicle.ini (values typed here as data). The reference reads
# the polygon as float (TrajectoriesAndRobotShape.cpp:25-27) and w_max as deg -> x*pi/180.
ACKERMANN_SHAPE_X = np.array([-0.10, 0.00, 0.3, 0.3, 0.0, -0.10], dtype=np.float32).astype(np.float64)
ACKERMANN_SHAPE_Y = np.array([0.15, 0.15, 0.1, -0.1, -0.15, -0.15], dtype=np.float32).astype(np.float64)


def ackermann_ptgs(num_paths=121, ref_distance=5.0, resolution=0.05):
    w = 60.0 * np.pi / 180.0
    return [PTG.diffdrive_c(num_paths, ref_distance, resolution, 1.0, w, K, ACKERMANN_SHAPE_X, ACKERMANN_SHAPE_Y)
            for K in (+1.0, -1.0)]
