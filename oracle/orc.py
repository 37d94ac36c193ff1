"""ORACLE — TEST INFRASTRUCTURE ONLY (ctypes binding of oracle/_build/liboracle.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module. The product (mrpt_path_planning_b200) never does.
Pinned against the reference build oracle/_ref (tests/test_oracle_vs_reference.py); MRPT-layer arithmetic assumed.
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
        L.orc_solve_poly4.argtypes = [C.c_double] * 4 + [c_double_p]
        L.orc_solve_poly3.argtypes = [C.c_double] * 3 + [c_double_p]
        L.orc_solve_poly2.argtypes = [C.c_double] * 3 + [c_double_p]
        L.orc_holo_new.restype = C.c_void_p
        L.orc_holo_new.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_char_p,
                                   C.c_char_p]
        L.orc_ptg_update_dyn_state.argtypes = [C.c_void_p, c_double_p, C.c_int]
        L.orc_ptg_set_trim_speed.argtypes = [C.c_void_p, C.c_double]
        L.orc_ptg_path_twist.argtypes = [C.c_void_p, C.c_int, C.c_uint, c_double_p]
        L.orc_ptg_update_tpobs_single.restype = C.c_double
        L.orc_ptg_update_tpobs_single.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_double]
        L.orc_eval_candidates.restype = C.c_double
        L.orc_eval_candidates.argtypes = [C.c_void_p, c_float_p, c_float_p, C.c_size_t, c_double_p, c_double_p,
                                          C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int), c_double_p, C.c_size_t,
                                          C.c_double, c_double_p, c_double_p]
        L.orc_costmap_new.restype = C.c_void_p
        L.orc_costmap_new.argtypes = [c_float_p, c_float_p, C.c_size_t, C.c_double, C.c_double, C.c_double, C.c_int,
                                      C.c_double, c_double_p]
        L.orc_costmap_free.argtypes = [C.c_void_p]
        L.orc_costmap_info.argtypes = [C.c_void_p, c_double_p]
        L.orc_costmap_cells.argtypes = [C.c_void_p, c_double_p]
        L.orc_costmap_eval_pose.restype = C.c_double
        L.orc_costmap_eval_pose.argtypes = [C.c_void_p, C.c_double, C.c_double]
        L.orc_waypoints_new.restype = C.c_void_p
        L.orc_waypoints_new.argtypes = [c_double_p, c_double_p, C.c_size_t, C.c_double, C.c_double, C.c_int]
        L.orc_waypoints_free.argtypes = [C.c_void_p]
        L.orc_waypoints_eval_pose.restype = C.c_double
        L.orc_waypoints_eval_pose.argtypes = [C.c_void_p, C.c_double, C.c_double]
        L.orc_eval_edge_costs.argtypes = [C.c_void_p, C.c_int, C.c_size_t, c_double_p, C.POINTER(C.c_uint32),
                                          c_double_p, c_double_p]
        L.orc_planner_new.restype = C.c_void_p
        L.orc_planner_free.argtypes = [C.c_void_p]
        L.orc_planner_set_params.argtypes = [C.c_void_p, c_double_p, c_double_p, C.c_int]
        L.orc_planner_add_ptg.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_planner_add_obstacles.argtypes = [C.c_void_p, c_float_p, c_float_p, C.c_size_t]
        L.orc_planner_add_costmap.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_planner_add_waypoints.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_planner_plan.argtypes = [C.c_void_p, c_double_p, c_double_p, C.c_int, c_double_p, c_double_p,
                                       c_double_p]
        L.orc_planner_refine.argtypes = [C.c_void_p]
        L.orc_planner_path_len.restype = C.c_long
        L.orc_planner_path_len.argtypes = [C.c_void_p]
        L.orc_planner_path.argtypes = [C.c_void_p, c_double_p, c_double_p]
        L.orc_planner_tree.restype = C.c_long
        L.orc_planner_tree.argtypes = [C.c_void_p, c_double_p, C.c_long]
        L.orc_planner_trajectory_txt.restype = C.c_long
        L.orc_planner_trajectory_txt.argtypes = [C.c_void_p, C.c_double, C.c_char_p, C.c_long]
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

    def update_dyn_state(self, vel_local=(0, 0, 0), rel_target=(20.0, 0, 0), target_rel_speed=0.0, force=False):
        d = _f64(list(vel_local) + list(rel_target) + [target_rel_speed])
        if lib().orc_ptg_update_dyn_state(self.h, _p(d, c_double_p), int(force)) != 0:
            raise RuntimeError("oracle: " + _err())

    def set_trim_speed(self, s):
        lib().orc_ptg_set_trim_speed(self.h, s)

    def path_twist(self, k, step):
        out = np.zeros(3)
        if lib().orc_ptg_path_twist(self.h, k, step, _p(out, c_double_p)) != 0:
            raise RuntimeError("oracle: " + _err())
        return out

    def update_tpobs_single(self, ox, oy, k, tp_obstacle_k):
        return lib().orc_ptg_update_tpobs_single(self.h, ox, oy, k, tp_obstacle_k)

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


def holonomic_blend(num_paths, ref_distance, T_ramp_max, v_max, w_max_rad, robot_radius, expr_V="", expr_W=""):
    return PTG(lib().orc_holo_new(num_paths, ref_distance, T_ramp_max, v_max, w_max_rad, robot_radius,
                                  expr_V.encode(), expr_W.encode()))


# share/ptgs_holonomic_robot.ini (values typed in here as data)
HOLO_EXPR_V = "V_MAX * trimmable_speed"
HOLO_EXPR_W = "W_MAX * trimmable_speed * min(1.0, 0.1+abs(dir)/(10*PI/180))"


def holonomic_ptgs():
    return [holonomic_blend(191, 5.0, 1.0, 1.0, 60.0 * np.pi / 180.0, 0.15, HOLO_EXPR_V, HOLO_EXPR_W)]


def solve_poly4(a, b, c, d):
    x = np.zeros(4)
    n = lib().orc_solve_poly4(a, b, c, d, _p(x, c_double_p))
    return n, x


def solve_poly3(a, b, c):
    x = np.zeros(3)
    n = lib().orc_solve_poly3(a, b, c, _p(x, c_double_p))
    return n, x


def solve_poly2(a, b, c):
    x = np.zeros(2)
    n = lib().orc_solve_poly2(a, b, c, _p(x, c_double_p))
    return n, x


def eval_candidates(ptg, xs, ys, nodes, rel_targets, cand_node, cand_k, cand_speed, clip, want_free=True):
    """Per-candidate (node, k, speed) raw min collision distance (+inf if none) and freeDistance."""
    xs, ys = _f32(xs), _f32(ys)
    nodes = _f64(nodes).reshape(-1, 6)
    rel_targets = _f64(rel_targets).reshape(-1, 3)
    cn = np.ascontiguousarray(cand_node, dtype=np.int32)
    ck = np.ascontiguousarray(cand_k, dtype=np.int32)
    cs = _f64(cand_speed)
    raw = np.zeros(len(cn))
    free = np.zeros(len(cn))
    t = lib().orc_eval_candidates(ptg.h, _p(xs, c_float_p), _p(ys, c_float_p), len(xs), _p(nodes, c_double_p),
                                  _p(rel_targets, c_double_p), len(nodes), _p(cn, C.POINTER(C.c_int)),
                                  _p(ck, C.POINTER(C.c_int)), _p(cs, c_double_p), len(cn), clip,
                                  _p(raw, c_double_p), _p(free, c_double_p) if want_free else None)
    if t < 0:
        raise RuntimeError("oracle: " + _err())
    return raw, free, t


class CostMap:
    def __init__(self, xs, ys, resolution, clearance, max_cost, use_average, max_radius_from_robot, robot_pose):
        xs, ys = _f32(xs), _f32(ys)
        rp = _f64(robot_pose) if robot_pose is not None else None
        h = lib().orc_costmap_new(_p(xs, c_float_p), _p(ys, c_float_p), len(xs), resolution, clearance, max_cost,
                                  int(use_average), max_radius_from_robot, _p(rp, c_double_p) if rp is not None else None)
        if not h:
            raise RuntimeError("oracle: " + _err())
        self.h = C.c_void_p(h)
        self.kind = 0

    def info(self):
        v = np.zeros(5)
        lib().orc_costmap_info(self.h, _p(v, c_double_p))
        return dict(xmin=v[0], ymin=v[1], res=v[2], nx=int(v[3]), ny=int(v[4]))

    def cells(self):
        i = self.info()
        out = np.zeros((i["ny"], i["nx"]))
        lib().orc_costmap_cells(self.h, _p(out, c_double_p))
        return out

    def eval_pose(self, x, y):
        return lib().orc_costmap_eval_pose(self.h, x, y)

    def __del__(self):
        try:
            lib().orc_costmap_free(self.h)
        except Exception:
            pass


class Waypoints:
    def __init__(self, xs, ys, radius, scale, use_average):
        xs, ys = _f64(xs), _f64(ys)
        h = lib().orc_waypoints_new(_p(xs, c_double_p), _p(ys, c_double_p), len(xs), radius, scale, int(use_average))
        if not h:
            raise RuntimeError("oracle: " + _err())
        self.h = C.c_void_p(h)
        self.kind = 1

    def eval_pose(self, x, y):
        return lib().orc_waypoints_eval_pose(self.h, x, y)

    def __del__(self):
        try:
            lib().orc_waypoints_free(self.h)
        except Exception:
            pass


def eval_edge_costs(evaluator, from_xyphi, sample_offsets, rel_xy):
    f = _f64(from_xyphi).reshape(-1, 3)
    off = np.ascontiguousarray(sample_offsets, dtype=np.uint32)
    rel = _f64(rel_xy).reshape(-1, 2)
    out = np.zeros(len(f))
    if lib().orc_eval_edge_costs(evaluator.h, evaluator.kind, len(f), _p(f, c_double_p),
                                 _p(off, C.POINTER(C.c_uint32)), _p(rel, c_double_p), _p(out, c_double_p)) != 0:
        raise RuntimeError("oracle: " + _err())
    return out


RESULT_KEYS = ["success", "path_cost", "tree_nodes", "tree_parents", "expanded", "tps_points", "edges_costed",
               "best_node", "goal_node", "time_s", "best_cost_to_goal"]


class Planner:
    """mpp::TPS_Astar restatement (oracle)."""

    def __init__(self, ptgs, params, timestamps):
        self.h = C.c_void_p(lib().orc_planner_new())
        self._keep = list(ptgs)
        p, ts = _f64(params), _f64(timestamps)
        lib().orc_planner_set_params(self.h, _p(p, c_double_p), _p(ts, c_double_p), len(ts))
        for g in ptgs:
            lib().orc_planner_add_ptg(self.h, g.h)

    def add_obstacles(self, xs, ys):
        xs, ys = _f32(xs), _f32(ys)
        lib().orc_planner_add_obstacles(self.h, _p(xs, c_float_p), _p(ys, c_float_p), len(xs))

    def add_costmap(self, cm):
        self._keep.append(cm)
        lib().orc_planner_add_costmap(self.h, cm.h)

    def add_waypoints(self, wp):
        self._keep.append(wp)
        lib().orc_planner_add_waypoints(self.h, wp.h)

    def plan(self, start, goal, bbox_min, bbox_max, goal_is_point=False):
        s, g, lo, hi = _f64(start), _f64(goal), _f64(bbox_min), _f64(bbox_max)
        res = np.zeros(11)
        if lib().orc_planner_plan(self.h, _p(s, c_double_p), _p(g, c_double_p), int(goal_is_point),
                                  _p(lo, c_double_p), _p(hi, c_double_p), _p(res, c_double_p)) != 0:
            raise RuntimeError("oracle: " + _err())
        return dict(zip(RESULT_KEYS, res.tolist()))

    def refine(self):
        if lib().orc_planner_refine(self.h) != 0:
            raise RuntimeError("oracle: " + _err())

    def path(self):
        n = lib().orc_planner_path_len(self.h)
        nodes = np.zeros((n, 8))
        edges = np.zeros((max(n - 1, 0), 8))
        if n and lib().orc_planner_path(self.h, _p(nodes, c_double_p), _p(edges, c_double_p)) != 0:
            raise RuntimeError("oracle: " + _err())
        return nodes, edges

    def tree(self):
        n = lib().orc_planner_tree(self.h, None, 0)
        out = np.zeros((n, 7))
        lib().orc_planner_tree(self.h, _p(out, c_double_p), n)
        return out

    def trajectory_txt(self, period=0.05):
        n = lib().orc_planner_trajectory_txt(self.h, period, None, 0)
        if n < 0:
            raise RuntimeError("oracle: " + _err())
        buf = C.create_string_buffer(n + 1)
        lib().orc_planner_trajectory_txt(self.h, period, buf, n + 1)
        return buf.value.decode()

    def __del__(self):
        try:
            lib().orc_planner_free(self.h)
        except Exception:
            pass
