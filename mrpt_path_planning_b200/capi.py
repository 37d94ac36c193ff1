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


class HoloDesc(C.Structure):
    _fields_ = [("num_paths", C.c_int32), ("ref_distance", C.c_double), ("robot_radius", C.c_double),
                ("v_max", C.c_double)]


class HoloCand(C.Structure):
    _fields_ = [("node", C.c_int32), ("ptg_index", C.c_int32), ("T_ramp", C.c_double), ("vxi", C.c_double),
                ("vyi", C.c_double), ("vxf", C.c_double), ("vyf", C.c_double), ("vf", C.c_double)]


HOLO_CAND_DTYPE = np.dtype([("node", np.int32), ("ptg_index", np.int32), ("T_ramp", np.float64),
                            ("vxi", np.float64), ("vyi", np.float64), ("vxf", np.float64), ("vyf", np.float64),
                            ("vf", np.float64)])
assert HOLO_CAND_DTYPE.itemsize == C.sizeof(HoloCand) == 56


class HoloParams(C.Structure):
    _fields_ = [("num_paths", C.c_int32), ("ref_distance", C.c_double), ("T_ramp_max", C.c_double),
                ("v_max_mps", C.c_double), ("w_max_rps", C.c_double), ("robot_radius", C.c_double),
                ("expr_V", C.c_char_p), ("expr_W", C.c_char_p)]


class CostmapDesc(C.Structure):
    _fields_ = [("x_min", C.c_double), ("y_min", C.c_double), ("res", C.c_double), ("nx", C.c_int32),
                ("ny", C.c_int32), ("cells", c_double_p), ("use_average", C.c_int32)]


class AstarParams(C.Structure):
    _fields_ = [("grid_resolution_xy", C.c_double), ("grid_resolution_yaw", C.c_double),
                ("SE2_metricAngleWeight", C.c_double), ("heuristic_heading_weight", C.c_double),
                ("max_ptg_trajectories_to_explore", C.c_uint32), ("max_ptg_speeds_to_explore", C.c_uint32),
                ("path_interpolated_segments", C.c_uint32), ("maximum_computation_time", C.c_double),
                ("max_expansions", C.c_uint64), ("ptg_sample_timestamps", c_double_p), ("num_timestamps", C.c_uint32)]


MAX_PTGS = 8
COMM_ID_BYTES = 128
IPC_HANDLE_BYTES = 64


def comm_unique_id():
    """ncclGetUniqueId through the library (call on one rank, hand the bytes to the others)."""
    buf = (C.c_uint8 * COMM_ID_BYTES)()
    if lib().mppb_comm_unique_id(buf) != 0:
        raise MppbError(lib().mppb_last_global_error().decode())
    return bytes(buf)



class PathsDesc(C.Structure):
    _fields_ = [("num_paths", C.c_int32), ("path_begin", c_u32_p), ("x", c_float_p), ("y", c_float_p),
                ("phi", c_float_p), ("dist", c_float_p), ("v", c_float_p), ("w", c_float_p), ("cos_phi", c_double_p),
                ("sin_phi", c_double_p), ("step_duration", C.c_double), ("ref_distance", C.c_double)]


EXPAND_NODE_DTYPE = np.dtype([("pose", np.float64, 3), ("bbox_min", np.float64, 3), ("bbox_max", np.float64, 3),
                              ("goal_k", np.int32, MAX_PTGS), ("goal_step", np.uint32, MAX_PTGS),
                              ("cos_phi", np.float64), ("sin_phi", np.float64)])
NEIGHBOR_DTYPE = np.dtype([("cell", np.int32, 3), ("ptg_index", np.uint8), ("in_bbox", np.uint8), ("k", np.uint16),
                           ("step", np.uint32), ("ptg_dist", np.float32), ("pose", np.float64, 3),
                           ("vel", np.float64, 3), ("cost", np.float64)])
assert EXPAND_NODE_DTYPE.itemsize == 152 and NEIGHBOR_DTYPE.itemsize == 80


class PlanQuery(C.Structure):
    _fields_ = [("start", C.c_double * 6), ("goal", C.c_double * 6), ("goal_is_point", C.c_int32),
                ("bbox_min", C.c_double * 3), ("bbox_max", C.c_double * 3)]


class PlanResult(C.Structure):
    _fields_ = [("success", C.c_int32), ("path_cost", C.c_double), ("tree_nodes", C.c_uint64),
                ("tree_parents", C.c_uint64), ("nodes_expanded", C.c_uint64), ("tps_points", C.c_uint64),
                ("edges_costed", C.c_uint64), ("best_node", C.c_int64), ("goal_node", C.c_int64),
                ("computation_time", C.c_double), ("best_cost_to_goal", C.c_double), ("path_len", C.c_uint32)]


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
    "mppb_obstacles_epoch": (C.c_uint64, [C.c_void_p]),
    "mppb_ptgs_epoch": (C.c_uint64, [C.c_void_p]),
    "mppb_ctx_device": (C.c_int, [C.c_void_p]),
    "mppb_clear_ptgs": (C.c_int, [C.c_void_p]),
    "mppb_set_ptg_grid": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(GridDesc)]),
    "mppb_num_ptgs": (C.c_int, [C.c_void_p]),
    "mppb_total_paths": (C.c_int, [C.c_void_p]),
    "mppb_eval_nodes": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_double, C.c_void_p]),
    "mppb_eval_nodes_device": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_double, C.c_void_p]),
    "mppb_nodes_xycs_from_poses": (None, [C.c_size_t, C.c_void_p, C.c_void_p]),
    "mppb_ptg_create_diffdrive_c": (C.c_int, [C.POINTER(DiffDriveCParams), C.POINTER(C.c_void_p)]),
    # (desc and result are passed as raw pointers: the host PTG object is the caller, see host/ptg.cpp)
    "mppb_build_collision_grid": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "mppb_colgrid_build_stats": (C.c_int, [C.c_void_p, c_double_p]),
    "mppb_ptg_create_diffdrive_c_on": (C.c_int, [C.c_void_p, C.POINTER(DiffDriveCParams), C.POINTER(C.c_void_p)]),
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
    "mppb_set_obstacles_clipped": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_double, C.c_int,
                                             C.c_void_p]),
    "mppb_ptg_column_offset": (C.c_int, [C.c_void_p, C.c_int]),
    "mppb_set_ptg_holo": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(HoloDesc)]),
    "mppb_eval_holo_candidates": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_double,
                                            C.c_void_p]),
    "mppb_eval_holo_candidates_device": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_double,
                                                   C.c_void_p]),
    "mppb_clear_evaluators": (C.c_int, [C.c_void_p]),
    "mppb_set_costmap": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(CostmapDesc)]),
    "mppb_set_waypoints": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_double,
                                     C.c_double, C.c_int]),
    "mppb_set_holo_filter": (C.c_int, [C.c_void_p, C.c_int]),
    "mppb_set_grid_near": (C.c_int, [C.c_void_p, C.c_int]),
    "mppb_grid_near_stats": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "mppb_grid_near_window": (C.c_int, [C.c_void_p, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "mppb_set_holo_by_node": (C.c_int, [C.c_void_p, C.c_int]),
    "mppb_evaluators_epoch": (C.c_uint64, [C.c_void_p]),
    "mppb_num_evaluators": (C.c_int, [C.c_void_p]),
    "mppb_eval_edge_costs": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mppb_eval_edge_costs_device": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p,
                                              C.c_void_p]),
    "mppb_costmap_nearest_d2": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_double, C.c_double,
                                          C.c_double, C.c_int, C.c_int, C.c_float, C.c_void_p]),
    "mppb_ptg_create_holo": (C.c_int, [C.POINTER(HoloParams), C.POINTER(C.c_void_p)]),
    "mppb_ptg_update_dyn_state": (C.c_int, [C.c_void_p, c_double_p, C.c_int]),
    "mppb_ptg_set_trim_speed": (C.c_int, [C.c_void_p, C.c_double]),
    "mppb_ptg_path_twist": (C.c_int, [C.c_void_p, C.c_int, C.c_uint32, c_double_p]),
    "mppb_ptg_holo_params": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(HoloCand)]),
    "mppb_eval_nodes_boxed": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]),
    "mppb_eval_nodes_boxed_device": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_double,
                                               C.c_void_p]),
    "mppb_eval_nodes_boxed_begin": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_double,
                                              C.c_void_p]),
    "mppb_eval_holo_candidates_boxed_begin": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t,
                                                        C.c_void_p, C.c_double, C.c_void_p]),
    "mppb_eval_holo_candidates_boxed": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t,
                                                  C.c_void_p, C.c_double, C.c_void_p]),
    "mppb_eval_holo_candidates_boxed_device": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p,
                                                         C.c_double, C.c_void_p]),
    "mppb_selftest_worker_pool": (C.c_int, [C.c_int, C.c_size_t, C.c_int]),
    "mppb_selftest_open_set": (C.c_int, [C.c_uint64, C.c_size_t, C.c_int]),
    "mppb_selftest_lattice_index": (C.c_int, [C.c_uint64, C.c_size_t, C.c_int]),
    "mppb_selftest_cell_order": (C.c_int, [C.c_uint64, C.c_size_t, C.c_int]),
    "mppb_selftest_sincos_helper": (C.c_int, [C.c_uint64, C.c_size_t, C.c_int]),
    "mppb_planner_create": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "mppb_planner_destroy": (None, [C.c_void_p]),
    "mppb_planner_last_error": (C.c_char_p, [C.c_void_p]),
    "mppb_planner_set_params": (C.c_int, [C.c_void_p, C.POINTER(AstarParams)]),
    "mppb_planner_add_ptg": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mppb_planner_add_obstacles": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "mppb_planner_update_obstacles": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t]),
    "mppb_planner_add_costmap": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_double, C.c_double,
                                           C.c_double, C.c_int, C.c_double, C.c_void_p]),
    "mppb_planner_costmap_cells": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "mppb_planner_add_waypoints": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_double, C.c_double,
                                             C.c_int]),
    "mppb_planner_plan": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(PlanQuery), C.c_int, C.POINTER(PlanResult)]),
    "mppb_planner_refine": (C.c_int, [C.c_void_p, C.c_size_t]),
    "mppb_planner_path": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "mppb_planner_tree": (C.c_int64, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_int64]),
    "mppb_planner_trajectory_txt": (C.c_int64, [C.c_void_p, C.c_size_t, C.c_double, C.c_char_p, C.c_int64]),
    "mppb_planner_gpu_stats": (C.c_int, [C.c_void_p, c_double_p]),
    "mppb_planner_plan_multi": (C.c_int, [C.c_void_p, C.c_int, C.c_size_t, C.c_void_p, C.c_int, C.c_void_p]),
    "mppb_transform_pc_square_clipping": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_double,
                                                  C.c_void_p, C.c_void_p, C.POINTER(C.c_size_t)]),
    "mppb_comm_unique_id": (C.c_int, [C.c_void_p]),
    "mppb_comm_init": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "mppb_comm_world": (C.c_int, [C.c_void_p]),
    "mppb_comm_rank": (C.c_int, [C.c_void_p]),
    "mppb_set_obstacles_shard": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_double, C.c_int, C.c_int]),
    "mppb_allreduce_min_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "mppb_peer_rows_create": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p]),
    "mppb_peer_rows_attach": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "mppb_eval_nodes_sharded_device": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_double, C.c_void_p]),
    "mppb_sharded_mode": (C.c_int, [C.c_void_p]),
    "mppb_set_ptg_paths": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(PathsDesc)]),
    "mppb_ptg_paths_export": (C.c_int, [C.c_void_p, C.POINTER(PathsDesc)]),
    "mppb_set_expand_params": (C.c_int, [C.c_void_p, C.POINTER(AstarParams)]),
    "mppb_expand_max_neighbors": (C.c_uint32, [C.c_void_p]),
    "mppb_expand_nodes": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_void_p]),
    "mppb_expand_nodes_begin": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_double, C.c_void_p,
                                          C.c_void_p]),
    "mppb_expand_last_rows": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p]),
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
                    turning_radius_ref=0.10, ctx=None):
        """ctx=None: collision grid from the host builder; ctx=Context: built on that GPU (same table)."""
        sx = np.ascontiguousarray(shape_x, dtype=np.float64)
        sy = np.ascontiguousarray(shape_y, dtype=np.float64)
        p = DiffDriveCParams(num_paths, ref_distance, resolution, v_max_mps, w_max_rps, K, turning_radius_ref,
                             len(sx), sx.ctypes.data_as(c_double_p), sy.ctypes.data_as(c_double_p))
        h = C.c_void_p()
        if lib().mppb_ptg_create_diffdrive_c_on(ctx.h if ctx is not None else None, C.byref(p), C.byref(h)) != 0:
            raise MppbError(lib().mppb_ptg_last_error().decode())
        return PTG(h.value)

    @staticmethod
    def holonomic_blend(num_paths, ref_distance, T_ramp_max, v_max_mps, w_max_rps, robot_radius, expr_V="",
                        expr_W=""):
        p = HoloParams(num_paths, ref_distance, T_ramp_max, v_max_mps, w_max_rps, robot_radius, expr_V.encode(),
                       expr_W.encode())
        h = C.c_void_p()
        if lib().mppb_ptg_create_holo(C.byref(p), C.byref(h)) != 0:
            raise MppbError(lib().mppb_ptg_last_error().decode())
        return PTG(h.value)

    def _ck(self, rc):
        if rc != 0:
            raise MppbError(lib().mppb_ptg_last_error().decode())

    def update_dyn_state(self, vel_local=(0, 0, 0), rel_target=(20.0, 0, 0), target_rel_speed=0.0, force=False):
        d = np.array(list(vel_local) + list(rel_target) + [target_rel_speed], dtype=np.float64)
        self._ck(lib().mppb_ptg_update_dyn_state(self.h, d.ctypes.data_as(c_double_p), int(force)))

    def set_trim_speed(self, s):
        self._ck(lib().mppb_ptg_set_trim_speed(self.h, s))

    def path_twist(self, k, step):
        out = np.zeros(3)
        self._ck(lib().mppb_ptg_path_twist(self.h, k, step, out.ctypes.data_as(c_double_p)))
        return out

    def holo_params(self, k):
        c = HoloCand()
        self._ck(lib().mppb_ptg_holo_params(self.h, k, C.byref(c)))
        return c

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

    def paths_desc(self):
        """mppb_ptg_paths_desc of a collision-grid PTG (pointers into this object's tables)"""
        d = PathsDesc()
        if lib().mppb_ptg_paths_export(self.h, C.byref(d)) != 0:
            raise MppbError(lib().mppb_ptg_last_error().decode())
        return d

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

    def set_holo_filter(self, on):
        """Diagnostic: tpobs_holo_kernel without its geometric pre-filter (0/False), with pre-filter and bin culling
        (1/True, the default), or with the pre-filter over the whole window (2)."""
        self._ck(lib().mppb_set_holo_filter(self.h, int(on)))

    def set_grid_near(self, mode):
        """Near-first evaluation of node batches (mppb_set_grid_near): 0 off, 1 adaptive (default), 2 always."""
        self._ck(lib().mppb_set_grid_near(self.h, int(mode)))

    def grid_near_stats(self):
        """(calls in the two-launch form, their nodes, nodes that needed the second launch)"""
        out = (C.c_uint64 * 3)()
        self._ck(lib().mppb_grid_near_stats(self.h, out))
        return int(out[0]), int(out[1]), int(out[2])

    def grid_near_window(self, clip_dist):
        """(half-size of the near window, bound) in metres for this clipping distance; (0, 0): not applicable"""
        r, b = C.c_double(0), C.c_double(0)
        self._ck(lib().mppb_grid_near_window(self.h, float(clip_dist), C.byref(r), C.byref(b)))
        return r.value, b.value

    def set_holo_by_node(self, on):
        """Diagnostic: grouped host batches on the one-CTA-per-node form of the HolonomicBlend kernel."""
        self._ck(lib().mppb_set_holo_by_node(self.h, int(on)))

    def colgrid_build_stats(self):
        """Timings of the last mppb_build_collision_grid on this context."""
        v = np.zeros(4)
        self._ck(lib().mppb_colgrid_build_stats(self.h, v.ctypes.data_as(c_double_p)))
        return {"host_sincos_s": v[0], "gpu_total_s": v[1], "kernels_s": v[2], "steps": int(v[3])}

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

    def set_obstacles_clipped(self, xs, ys, box, bin_size=0.0, append=False):
        xs = np.ascontiguousarray(xs, dtype=np.float32)
        ys = np.ascontiguousarray(ys, dtype=np.float32)
        b = np.ascontiguousarray(box, dtype=np.float32) if box is not None else None
        self._ck(lib().mppb_set_obstacles_clipped(self.h, _ptr(xs), _ptr(ys), xs.size, bin_size, int(append),
                                                  _ptr(b)))

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

    def column_offset(self, ptg_index):
        return lib().mppb_ptg_column_offset(self.h, ptg_index)

    def eval_holo_candidates(self, poses, cands, clip_dist):
        """poses [n][3]; cands: numpy structured array (HOLO_CAND_DTYPE) -> float64 [len(cands)]."""
        poses = np.ascontiguousarray(poses, dtype=np.float64).reshape(-1, 3)
        cands = np.ascontiguousarray(cands, dtype=HOLO_CAND_DTYPE)
        out = np.empty(len(cands), dtype=np.float64)
        self._ck(lib().mppb_eval_holo_candidates(self.h, poses.shape[0], _ptr(poses), len(cands), _ptr(cands),
                                                 clip_dist, _ptr(out)))
        return out

    def eval_holo_candidates_device(self, n_cands, d_cands, d_nodes_xycs, clip_dist, d_out):
        self._ck(lib().mppb_eval_holo_candidates_device(self.h, n_cands, _ptr(d_cands), _ptr(d_nodes_xycs),
                                                        clip_dist, _ptr(d_out)))

    # -- free-function seams --
    def transform_pc_square_clipping(self, xs, ys, pose, max_dist):
        """mpp::transform_pc_square_clipping for one pose -> (local_x, local_y) float32, input order"""
        xs = np.ascontiguousarray(xs, dtype=np.float32)
        ys = np.ascontiguousarray(ys, dtype=np.float32)
        pose = np.ascontiguousarray(pose, dtype=np.float64)
        ox, oy = np.empty_like(xs), np.empty_like(ys)
        n = C.c_size_t(0)
        self._ck(lib().mppb_transform_pc_square_clipping(self.h, _ptr(xs), _ptr(ys), xs.size, _ptr(pose), max_dist, _ptr(ox),
                                                         _ptr(oy), C.byref(n)))
        return ox[:n.value].copy(), oy[:n.value].copy()

    # -- several GPUs: cloud sharding + min-merge --
    def set_obstacles_shard(self, xs, ys, shard, n_shards, bin_size=0.0):
        xs = np.ascontiguousarray(xs, dtype=np.float32)
        ys = np.ascontiguousarray(ys, dtype=np.float32)
        self._ck(lib().mppb_set_obstacles_shard(self.h, _ptr(xs), _ptr(ys), xs.size, bin_size, shard, n_shards))

    def comm_init(self, world, rank, unique_id):
        buf = (C.c_uint8 * COMM_ID_BYTES).from_buffer_copy(bytes(unique_id))
        self._ck(lib().mppb_comm_init(self.h, world, rank, buf))

    def peer_rows_create(self, max_nodes):
        buf = (C.c_uint8 * IPC_HANDLE_BYTES)()
        self._ck(lib().mppb_peer_rows_create(self.h, max_nodes, buf))
        return bytes(buf)

    def peer_rows_attach(self, world, rank, handles):
        blob = b"".join(bytes(h) for h in handles)
        assert len(blob) == world * IPC_HANDLE_BYTES
        buf = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
        self._ck(lib().mppb_peer_rows_attach(self.h, world, rank, buf))

    def allreduce_min_device(self, d_rows, count):
        self._ck(lib().mppb_allreduce_min_device(self.h, _ptr(d_rows), count))

    def eval_nodes_sharded_device(self, n_nodes, d_nodes_xycs, clip_dist, d_out):
        self._ck(lib().mppb_eval_nodes_sharded_device(self.h, n_nodes, _ptr(d_nodes_xycs), clip_dist, _ptr(d_out)))

    @property
    def sharded_mode(self):
        return {0: "plain", 1: "nccl", 2: "fused"}[lib().mppb_sharded_mode(self.h)]

    # -- one A* expansion per node on the device --
    def set_expand_params(self, params, timestamps):
        """params: the 9-list of workloads.DEMO_ASTAR_PARAMS (grid xy, grid yaw, ..., trajectories, speeds, segments)"""
        ts = np.ascontiguousarray(timestamps, dtype=np.float64)
        a = AstarParams(params[0], params[1], params[2], params[3], int(params[4]), int(params[5]), int(params[6]),
                        0.0, 0, ts.ctypes.data_as(c_double_p), len(ts))
        self._ck(lib().mppb_set_expand_params(self.h, C.byref(a)))

    @property
    def expand_max_neighbors(self):
        return lib().mppb_expand_max_neighbors(self.h)

    def expand_nodes(self, nodes, clip_dist, use_node_boxes=False):
        """nodes: structured array (EXPAND_NODE_DTYPE) -> (neighbors [n][stride] NEIGHBOR_DTYPE, counts [n][2] uint32)"""
        nodes = np.ascontiguousarray(nodes, dtype=EXPAND_NODE_DTYPE)
        stride = self.expand_max_neighbors
        if stride == 0:
            raise MppbError("device expansion not available: " + lib().mppb_last_error(self.h).decode())
        out = np.zeros((len(nodes), stride), dtype=NEIGHBOR_DTYPE)
        counts = np.zeros((len(nodes), 2), dtype=np.uint32)
        self._ck(lib().mppb_expand_nodes(self.h, len(nodes), _ptr(nodes), int(use_node_boxes), clip_dist, _ptr(out),
                                         _ptr(counts)))
        return out, counts

    def expand_last_rows(self, n):
        out = np.empty((n, self.total_paths), dtype=np.float32)
        self._ck(lib().mppb_expand_last_rows(self.h, n, _ptr(out)))
        return out

    # -- stage (c) --
    def clear_evaluators(self):
        self._ck(lib().mppb_clear_evaluators(self.h))

    def set_costmap(self, slot, cells, x_min, y_min, res, use_average):
        cells = np.ascontiguousarray(cells, dtype=np.float64)
        ny, nx = cells.shape
        d = CostmapDesc(x_min, y_min, res, nx, ny, cells.ctypes.data_as(c_double_p), int(use_average))
        self._ck(lib().mppb_set_costmap(self.h, slot, C.byref(d)))

    def set_waypoints(self, slot, xs, ys, radius, scale, use_average):
        xs = np.ascontiguousarray(xs, dtype=np.float64)
        ys = np.ascontiguousarray(ys, dtype=np.float64)
        self._ck(lib().mppb_set_waypoints(self.h, slot, _ptr(xs), _ptr(ys), xs.size, radius, scale, int(use_average)))

    @property
    def num_evaluators(self):
        return lib().mppb_num_evaluators(self.h)

    def eval_edge_costs(self, from_xyphi, sample_offsets, rel_xy):
        f = np.ascontiguousarray(from_xyphi, dtype=np.float64).reshape(-1, 3)
        off = np.ascontiguousarray(sample_offsets, dtype=np.uint32)
        rel = np.ascontiguousarray(rel_xy, dtype=np.float64).reshape(-1, 2)
        out = np.empty((len(f), self.num_evaluators), dtype=np.float64)
        self._ck(lib().mppb_eval_edge_costs(self.h, len(f), _ptr(f), _ptr(off), _ptr(rel), _ptr(out)))
        return out

    def costmap_nearest_d2(self, xs, ys, x_min, y_min, res, nx, ny, clearance):
        xs = np.ascontiguousarray(xs, dtype=np.float32)
        ys = np.ascontiguousarray(ys, dtype=np.float32)
        out = np.empty((ny, nx), dtype=np.float32)
        self._ck(lib().mppb_costmap_nearest_d2(self.h, _ptr(xs), _ptr(ys), xs.size, x_min, y_min, res, nx, ny,
                                               clearance, _ptr(out)))
        return out

    def eval_nodes_boxed(self, poses, boxes, clip_dist):
        poses = np.ascontiguousarray(poses, dtype=np.float64).reshape(-1, 3)
        boxes = np.ascontiguousarray(boxes, dtype=np.float32).reshape(-1, 4)
        out = np.empty((poses.shape[0], self.total_paths), dtype=np.float32)
        self._ck(lib().mppb_eval_nodes_boxed(self.h, poses.shape[0], _ptr(poses), _ptr(boxes), clip_dist, _ptr(out)))
        return out

    def eval_holo_candidates_boxed(self, poses, boxes, cands, clip_dist):
        poses = np.ascontiguousarray(poses, dtype=np.float64).reshape(-1, 3)
        boxes = np.ascontiguousarray(boxes, dtype=np.float32).reshape(-1, 4)
        cands = np.ascontiguousarray(cands, dtype=HOLO_CAND_DTYPE)
        out = np.empty(len(cands), dtype=np.float64)
        self._ck(lib().mppb_eval_holo_candidates_boxed(self.h, poses.shape[0], _ptr(poses), _ptr(boxes), len(cands),
                                                       _ptr(cands), clip_dist, _ptr(out)))
        return out

    def eval_both_split(self, poses, boxes, cands, clip_dist):
        """Collision-grid rows and HolonomicBlend candidates of the same nodes through the split entry points
        (mppb_eval_nodes_boxed_begin + mppb_eval_holo_candidates_boxed_begin, both in flight, one mppb_sync).
        boxes may be None; either result is None when the context has no PTG of that kind / cands is empty."""
        poses = np.ascontiguousarray(poses, dtype=np.float64).reshape(-1, 3)
        if boxes is not None:
            boxes = np.ascontiguousarray(boxes, dtype=np.float32).reshape(-1, 4)
        rows = hout = None
        if self.total_paths > 0:
            rows = np.empty((poses.shape[0], self.total_paths), dtype=np.float32)
            self._ck(lib().mppb_eval_nodes_boxed_begin(self.h, poses.shape[0], _ptr(poses),
                                                       _ptr(boxes) if boxes is not None else None, clip_dist,
                                                       _ptr(rows)))
        if cands is not None and len(cands):
            cands = np.ascontiguousarray(cands, dtype=HOLO_CAND_DTYPE)
            hout = np.empty(len(cands), dtype=np.float64)
            self._ck(lib().mppb_eval_holo_candidates_boxed_begin(self.h, poses.shape[0], _ptr(poses),
                                                                 _ptr(boxes) if boxes is not None else None,
                                                                 len(cands), _ptr(cands), clip_dist, _ptr(hout)))
        self.sync()
        return rows, hout

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


class _PinnedBlock:
    """Owner of one mppb_host_alloc block: exposes it through the array interface (numpy keeps this object as the
    base of every view) and returns it with mppb_host_free when the last view is collected."""

    def __init__(self, nbytes):
        self._p = C.c_void_p()
        if lib().mppb_host_alloc(nbytes, C.byref(self._p)) != 0:
            raise MppbError(lib().mppb_last_global_error().decode())
        self._free = lib().mppb_host_free  # bound now: module globals may be gone at interpreter shutdown
        self.__array_interface__ = {"shape": (max(1, nbytes),), "typestr": "|u1", "data": (self._p.value, False),
                                    "version": 3}

    def __del__(self):
        p, self._p = self._p, None
        if p is not None and p.value:
            self._free(p)


def pinned_empty(shape, dtype):
    """numpy array over pinned host memory from mppb_host_alloc (freed when the array and its views are collected)."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape))
    block = _PinnedBlock(n * dtype.itemsize)
    return np.asarray(block)[:n * dtype.itemsize].view(dtype).reshape(shape)


# PTG set of share/ptgs_ackermann_vehicle.ini (values typed here as data). The reference reads
# the polygon as float (TrajectoriesAndRobotShape.cpp:25-27) and w_max as deg -> x*pi/180.
ACKERMANN_SHAPE_X = np.array([-0.10, 0.00, 0.3, 0.3, 0.0, -0.10], dtype=np.float32).astype(np.float64)
ACKERMANN_SHAPE_Y = np.array([0.15, 0.15, 0.1, -0.1, -0.15, -0.15], dtype=np.float32).astype(np.float64)


def ackermann_ptgs(num_paths=121, ref_distance=5.0, resolution=0.05, ctx=None):
    w = 60.0 * np.pi / 180.0
    return [PTG.diffdrive_c(num_paths, ref_distance, resolution, 1.0, w, K, ACKERMANN_SHAPE_X, ACKERMANN_SHAPE_Y,
                            ctx=ctx)
            for K in (+1.0, -1.0)]


HOLO_EXPR_V = "V_MAX * trimmable_speed"
HOLO_EXPR_W = "W_MAX * trimmable_speed * min(1.0, 0.1+abs(dir)/(10*PI/180))"


def holonomic_ptgs():
    """PTG set of share/ptgs_holonomic_robot.ini (values typed here as data)."""
    return [PTG.holonomic_blend(191, 5.0, 1.0, 1.0, 60.0 * np.pi / 180.0, 0.15, HOLO_EXPR_V, HOLO_EXPR_W)]


RESULT_KEYS = ["success", "path_cost", "tree_nodes", "tree_parents", "expanded", "tps_points", "edges_costed",
               "best_node", "goal_node", "time_s", "best_cost_to_goal", "path_len"]


class Planner:
    """mpp::TPS_Astar over a libmpp_b200 context (C ABI section "planner")."""

    def __init__(self, ctx, ptgs, params, timestamps):
        """params: [grid_xy, grid_yaw_rad, SE2_metricAngleWeight, heuristic_heading_weight, max_trajs, max_speeds,
        path_interpolated_segments, max_time_s (<=0 unlimited), max_expansions (0 unlimited)]"""
        self.ctx = ctx
        self._keep = list(ptgs)
        h = C.c_void_p()
        if lib().mppb_planner_create(ctx.h, C.byref(h)) != 0:
            raise MppbError(lib().mppb_planner_last_error(None).decode())
        self.h = h
        ts = np.ascontiguousarray(timestamps, dtype=np.float64)
        a = AstarParams(params[0], params[1], params[2], params[3], int(params[4]), int(params[5]), int(params[6]),
                        params[7], int(params[8]), ts.ctypes.data_as(c_double_p), len(ts))
        self._ck(lib().mppb_planner_set_params(self.h, C.byref(a)))
        for p in ptgs:
            self._ck(lib().mppb_planner_add_ptg(self.h, p.h))

    def _ck(self, rc):
        if rc != 0:
            raise MppbError(lib().mppb_planner_last_error(self.h).decode())

    def close(self):
        if self.h:
            lib().mppb_planner_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def add_obstacles(self, xs, ys):
        xs = np.ascontiguousarray(xs, dtype=np.float32)
        ys = np.ascontiguousarray(ys, dtype=np.float32)
        self._ck(lib().mppb_planner_add_obstacles(self.h, _ptr(xs), _ptr(ys), xs.size))

    def update_obstacles(self, source, xs, ys):
        xs = np.ascontiguousarray(xs, dtype=np.float32)
        ys = np.ascontiguousarray(ys, dtype=np.float32)
        self._ck(lib().mppb_planner_update_obstacles(self.h, source, _ptr(xs), _ptr(ys), xs.size))

    def add_costmap(self, xs, ys, resolution, clearance, max_cost, use_average, max_radius_from_robot, robot_pose):
        xs = np.ascontiguousarray(xs, dtype=np.float32)
        ys = np.ascontiguousarray(ys, dtype=np.float32)
        rp = np.ascontiguousarray(robot_pose, dtype=np.float64) if robot_pose is not None else None
        self._ck(lib().mppb_planner_add_costmap(self.h, _ptr(xs), _ptr(ys), xs.size, resolution, clearance, max_cost,
                                                int(use_average), max_radius_from_robot, _ptr(rp)))

    def costmap_cells(self, evaluator=0):
        info = np.zeros(5)
        self._ck(lib().mppb_planner_costmap_cells(self.h, evaluator, _ptr(info), None))
        cells = np.zeros((int(info[4]), int(info[3])))
        self._ck(lib().mppb_planner_costmap_cells(self.h, evaluator, None, _ptr(cells)))
        return dict(xmin=info[0], ymin=info[1], res=info[2], nx=int(info[3]), ny=int(info[4])), cells

    def add_waypoints(self, xs, ys, radius, scale, use_average):
        xs = np.ascontiguousarray(xs, dtype=np.float64)
        ys = np.ascontiguousarray(ys, dtype=np.float64)
        self._ck(lib().mppb_planner_add_waypoints(self.h, _ptr(xs), _ptr(ys), xs.size, radius, scale,
                                                  int(use_average)))

    def plan_many(self, queries, n_threads=0):
        """queries: list of (start6, goal6, bbox_min3, bbox_max3, goal_is_point). Returns a list of dicts."""
        n = len(queries)
        qs = (PlanQuery * n)()
        for i, (start, goal, lo, hi, is_pt) in enumerate(queries):
            qs[i].start[:] = list(start)
            qs[i].goal[:] = list(goal)
            qs[i].goal_is_point = int(is_pt)
            qs[i].bbox_min[:] = list(lo)
            qs[i].bbox_max[:] = list(hi)
        rs = (PlanResult * n)()
        self._ck(lib().mppb_planner_plan(self.h, n, qs, n_threads, rs))
        out = []
        for r in rs:
            out.append(dict(success=float(r.success), path_cost=r.path_cost, tree_nodes=float(r.tree_nodes),
                            tree_parents=float(r.tree_parents), expanded=float(r.nodes_expanded),
                            tps_points=float(r.tps_points), edges_costed=float(r.edges_costed),
                            best_node=float(r.best_node), goal_node=float(r.goal_node), time_s=r.computation_time,
                            best_cost_to_goal=r.best_cost_to_goal, path_len=int(r.path_len)))
        self._path_len = [d["path_len"] for d in out]
        return out

    def plan(self, start, goal, bbox_min, bbox_max, goal_is_point=False):
        return self.plan_many([(start, goal, bbox_min, bbox_max, goal_is_point)])[0]

    @staticmethod
    def plan_multi(planners, queries, n_threads_per_planner=0):
        """mppb_planner_plan_multi: queries dealt round-robin to planners on several devices of this process; results
        in query order (query q is local query q // len(planners) of planners[q % len(planners)])."""
        n = len(queries)
        qs = (PlanQuery * n)()
        for i, (start, goal, lo, hi, is_pt) in enumerate(queries):
            qs[i].start[:] = list(start)
            qs[i].goal[:] = list(goal)
            qs[i].goal_is_point = int(is_pt)
            qs[i].bbox_min[:] = list(lo)
            qs[i].bbox_max[:] = list(hi)
        rs = (PlanResult * n)()
        hs = (C.c_void_p * len(planners))(*[p.h for p in planners])
        rc = lib().mppb_planner_plan_multi(hs, len(planners), n, qs, n_threads_per_planner, rs)
        if rc != 0:
            raise MppbError("; ".join(lib().mppb_planner_last_error(p.h).decode() for p in planners))
        out = [dict(success=float(r.success), path_cost=r.path_cost, tree_nodes=float(r.tree_nodes),
                    tree_parents=float(r.tree_parents), expanded=float(r.nodes_expanded), tps_points=float(r.tps_points),
                    edges_costed=float(r.edges_costed), best_node=float(r.best_node), goal_node=float(r.goal_node),
                    time_s=r.computation_time, best_cost_to_goal=r.best_cost_to_goal, path_len=int(r.path_len)) for r in rs]
        for i, p in enumerate(planners):
            p._path_len = [d["path_len"] for d in out[i::len(planners)]]
        return out

    def refine(self, query=0):
        self._ck(lib().mppb_planner_refine(self.h, query))

    def path(self, query=0):
        n = self._path_len[query]
        nodes = np.zeros((n, 8))
        edges = np.zeros((max(n - 1, 0), 8))
        self._ck(lib().mppb_planner_path(self.h, query, _ptr(nodes), _ptr(edges)))
        return nodes, edges

    def tree(self, query=0):
        n = lib().mppb_planner_tree(self.h, query, None, 0)
        if n < 0:
            raise MppbError(lib().mppb_planner_last_error(self.h).decode())
        out = np.zeros((n, 7))
        lib().mppb_planner_tree(self.h, query, _ptr(out), n)
        return out

    def trajectory_txt(self, query=0, period=0.05):
        n = lib().mppb_planner_trajectory_txt(self.h, query, period, None, 0)
        if n < 0:
            raise MppbError(lib().mppb_planner_last_error(self.h).decode())
        buf = C.create_string_buffer(n + 1)
        lib().mppb_planner_trajectory_txt(self.h, query, period, buf, n + 1)
        return buf.value.decode()

    def gpu_stats(self):
        v = np.zeros(14)
        self._ck(lib().mppb_planner_gpu_stats(self.h, v.ctypes.data_as(c_double_p)))
        names = ["setup_uploads", "F_relax+pop+A_enumerate", "B_collision_gpu", "CD_select_build", "E_assemble",
                 "E_cost_gpu", "teardown"]
        return dict(collision_batches=int(v[0]), cost_batches=int(v[1]), nodes=int(v[2]), holo_candidates=int(v[3]),
                    edges=int(v[4]), gpu_seconds=v[5], holo_near_ties=int(v[13]), phase_seconds={n: float(x) for n, x in zip(names, v[6:13]) if not n.startswith("_")})
