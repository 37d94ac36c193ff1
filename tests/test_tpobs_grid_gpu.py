"""GPU parity of stage (a)+(b) for collision-grid PTGs: libmpp_b200 (through the C ABI) vs the
CPU oracle on the same seeded inputs. Bar: bit-exact (the values are table floats, 0 or +inf)."""
import os

import numpy as np
import pytest

from conftest import free_dist
from mrpt_path_planning_b200 import workloads as wl

pytestmark = pytest.mark.gpu


def _check(gpu_ctx, orc, oracle_ptgs, xs, ys, poses, clip=5.0, bin_size=0.0):
    gpu_ctx.set_obstacles(xs, ys, bin_size=bin_size)
    out = gpu_ctx.eval_nodes(poses, clip)
    got = free_dist(out, oracle_ptgs)
    ref, _ = orc.eval_nodes(oracle_ptgs, xs, ys, poses, clip, mode=1, nthreads=0)
    assert got.shape == ref.shape
    bad = np.argwhere(got != ref)
    assert bad.size == 0, "first mismatches (node, col): %s got %s ref %s" % (
        bad[:5].tolist(), got[tuple(bad[:5].T)], ref[tuple(bad[:5].T)])
    return got


@pytest.mark.parametrize("density,seed", [(0.05, 1), (0.5, 2), (3.0, 3), (26.0, 4)])
def test_uniform_cloud_parity(gpu_ctx, orc, oracle_ackermann, density, seed):
    extent = 60.0
    n = int(density * extent * extent)
    xs, ys = wl.uniform_cloud(n, extent, seed=seed)
    poses = wl.node_poses(192, extent, margin=2.0, seed=seed + 100)
    got = _check(gpu_ctx, orc, oracle_ackermann, xs, ys, poses)
    if density <= 0.5:
        assert (got == 5.0).any() and (got < 5.0).any()  # both free and constrained paths exercised


def test_walls_cloud_parity(gpu_ctx, orc, oracle_ackermann):
    xs, ys = wl.walls_cloud(60000, 80.0, seed=11)
    poses = wl.node_poses(256, 80.0, margin=1.0, seed=12)
    _check(gpu_ctx, orc, oracle_ackermann, xs, ys, poses)


def test_reference_call_pattern_equals_all_k(orc, oracle_ackermann):
    """mode 0 (tp_obstacles_single_path per k) and mode 1 (all-k virtual) of the oracle agree."""
    xs, ys = wl.uniform_cloud(5000, 40.0, seed=5)
    poses = wl.node_poses(16, 40.0, seed=6)
    a, _ = orc.eval_nodes(oracle_ackermann, xs, ys, poses, 5.0, mode=0)
    b, _ = orc.eval_nodes(oracle_ackermann, xs, ys, poses, 5.0, mode=1)
    assert np.array_equal(a, b)


def test_empty_cloud(gpu_ctx, oracle_ackermann):
    gpu_ctx.set_obstacles(np.zeros(0, np.float32), np.zeros(0, np.float32))
    out = gpu_ctx.eval_nodes(wl.node_poses(7, 10.0, margin=1.0, seed=3), 5.0)
    assert out.shape == (7, 242) and np.isinf(out).all()


def test_zero_nodes(gpu_ctx):
    xs, ys = wl.uniform_cloud(100, 10.0, seed=1)
    gpu_ctx.set_obstacles(xs, ys)
    out = gpu_ctx.eval_nodes(np.zeros((0, 3)), 5.0)
    assert out.shape == (0, 242)


def test_nodes_outside_cloud_bbox(gpu_ctx, orc, oracle_ackermann):
    xs, ys = wl.uniform_cloud(3000, 20.0, seed=21)
    poses = np.array([[-100.0, -100.0, 0.3], [-4.0, 10.0, 1.0], [24.9, 10.0, -2.0], [10.0, 25.1, 3.0],
                      [1e6, 1e6, 0.0], [10.0, -4.99, 0.5], [10.0, 10.0, -3.14159]])
    _check(gpu_ctx, orc, oracle_ackermann, xs, ys, poses)


def test_nonfinite_points_are_ignored(gpu_ctx, orc, oracle_ackermann):
    xs, ys = wl.uniform_cloud(4000, 30.0, seed=31)
    xs[::97] = np.nan
    ys[5::101] = np.inf
    xs[7::89] = -np.inf
    poses = wl.node_poses(64, 30.0, margin=1.0, seed=32)
    _check(gpu_ctx, orc, oracle_ackermann, xs, ys, poses)


def test_clip_boundary_points(gpu_ctx, orc, oracle_ackermann):
    """Points exactly on, one ulp inside and one ulp outside the +-clip square (global axes)."""
    x0, y0 = 12.3, 45.6
    pts = []
    for base, other in ((x0 + 5.0, y0 + 0.2), (x0 - 5.0, y0 - 0.3)):
        f = np.float32(base)
        for v in (f, np.nextafter(f, np.float32(1e9)), np.nextafter(f, np.float32(-1e9))):
            pts.append((v, np.float32(other)))
            pts.append((np.float32(other - y0 + x0), np.float32(v - x0 + y0)))
    xs = np.array([p[0] for p in pts], np.float32)
    ys = np.array([p[1] for p in pts], np.float32)
    poses = np.array([[x0, y0, a] for a in np.linspace(-3.1, 3.1, 23)])
    _check(gpu_ctx, orc, oracle_ackermann, xs, ys, poses)


def test_points_near_robot(gpu_ctx, orc, oracle_ackermann):
    """Obstacles inside / on the edge of the robot polygon: the BACK_AWAY post-process branch."""
    rng = np.random.default_rng(41)
    xs = rng.uniform(9.5, 10.5, 400).astype(np.float32)
    ys = rng.uniform(9.5, 10.5, 400).astype(np.float32)
    for i in range(0, 400, 25):  # one point at a time: each node sees a single near obstacle
        poses = np.array([[10.0, 10.0, a] for a in np.linspace(-np.pi, np.pi, 9)])
        _check(gpu_ctx, orc, oracle_ackermann, xs[i:i + 1], ys[i:i + 1], poses)
    poses = np.stack([rng.uniform(9.0, 11.0, 64), rng.uniform(9.0, 11.0, 64), rng.uniform(-np.pi, np.pi, 64)], 1)
    _check(gpu_ctx, orc, oracle_ackermann, xs[:40], ys[:40], poses)


@pytest.mark.parametrize("bin_size", [0.1, 0.5, 3.0, 50.0])
def test_bin_size_independence(gpu_ctx, orc, oracle_ackermann, bin_size):
    xs, ys = wl.uniform_cloud(20000, 50.0, seed=51)
    poses = wl.node_poses(96, 50.0, margin=0.0, seed=52)
    _check(gpu_ctx, orc, oracle_ackermann, xs, ys, poses, bin_size=bin_size)


def test_other_clip_distance(gpu_ctx, orc, oracle_ackermann):
    xs, ys = wl.uniform_cloud(20000, 50.0, seed=61)
    poses = wl.node_poses(64, 50.0, margin=0.0, seed=62)
    for clip in (0.0, 1.25, 3.0, 7.5):
        _check(gpu_ctx, orc, oracle_ackermann, xs, ys, poses, clip=clip)


def test_append_sources(gpu_ctx, orc, oracle_ackermann):
    xs, ys = wl.uniform_cloud(9000, 40.0, seed=71)
    poses = wl.node_poses(64, 40.0, margin=0.0, seed=72)
    gpu_ctx.set_obstacles(xs[:4000], ys[:4000])
    gpu_ctx.set_obstacles(xs[4000:], ys[4000:], append=True)
    assert gpu_ctx.num_obstacles == 9000
    got = free_dist(gpu_ctx.eval_nodes(poses, 5.0), oracle_ackermann)
    ref, _ = orc.eval_nodes(oracle_ackermann, xs, ys, poses, 5.0, mode=1, nthreads=0)
    assert np.array_equal(got, ref)


def test_full_size_properties(gpu_ctx, oracle_ackermann):
    """Config C3 at full size (1M points, 4096 nodes): size-independent properties.
    (i) batch evaluation == evaluation of a sub-batch; (ii) min over a split cloud == whole cloud
    (the cloud-sharding identity of config C5); (iii) value set is {0} U table floats U {inf}."""
    xs, ys = wl.uniform_cloud(1 << 20, 200.0)
    poses = wl.node_poses(4096, 200.0)
    gpu_ctx.set_obstacles(xs, ys)
    full = gpu_ctx.eval_nodes(poses, 5.0)
    sub = gpu_ctx.eval_nodes(poses[1000:1100], 5.0)
    assert np.array_equal(full[1000:1100], sub)
    half = len(xs) // 2
    gpu_ctx.set_obstacles(xs[:half], ys[:half])
    a = gpu_ctx.eval_nodes(poses, 5.0)
    gpu_ctx.set_obstacles(xs[half:], ys[half:])
    b = gpu_ctx.eval_nodes(poses, 5.0)
    assert np.array_equal(np.minimum(a, b), full)
    assert (full >= 0).all() and not np.isnan(full).any()


def test_mixed_geometry_ptg_groups(capi, orc):
    """PTGs with different grid geometry / path counts form separate kernel groups."""
    from oracle.orc import ACKERMANN_SHAPE_X as X, ACKERMANN_SHAPE_Y as Y

    w = 60.0 * np.pi / 180.0
    # the last one has a 47 x 47 grid: partial 8 x 4 tiles of the occupancy bitmap on both axes
    specs = [(121, 5.0, 0.05, 1.0, w, +1.0), (31, 3.0, 0.1, 0.8, 1.3 * w, -1.0), (61, 5.0, 0.05, 1.0, w, -1.0),
             (25, 2.35, 0.1, 0.6, 1.1 * w, +1.0)]
    optgs = [orc.diffdrive_c(n, rd, res, v, ww, K, X, Y) for (n, rd, res, v, ww, K) in specs]
    pptgs = [capi.PTG.diffdrive_c(n, rd, res, v, ww, K, X, Y) for (n, rd, res, v, ww, K) in specs]
    ctx = capi.Context(0)
    for p in pptgs:
        ctx.add_ptg(p)
    assert ctx.total_paths == 121 + 31 + 61 + 25
    for density, seed in ((0.3, 81), (8.0, 82)):
        xs, ys = wl.uniform_cloud(int(density * 50 * 50), 50.0, seed=seed)
        poses = wl.node_poses(128, 50.0, margin=0.0, seed=seed + 10)
        _check(ctx, orc, optgs, xs, ys, poses, clip=5.0)
    ctx.close()


@pytest.mark.parametrize("density", [1.6, 2.1, 2.4, 2.7, 3.2, 6.0])
def test_sparse_dense_phase2_boundary(gpu_ctx, orc, oracle_ackermann, density):
    """Nodes on both sides of the point-count threshold (256 points in the bin window) that selects the scatter form
    or the column-scan form of phase 2 give the reference's values."""
    extent = 40.0
    xs, ys = wl.uniform_cloud(int(density * extent * extent), extent, seed=int(10 * density))
    poses = wl.node_poses(160, extent, margin=0.5, seed=int(10 * density) + 1)
    _check(gpu_ctx, orc, oracle_ackermann, xs, ys, poses)


def test_sparse_cloud_with_points_near_origin(gpu_ctx, orc, oracle_ackermann):
    """Sparse form with many long by-cell lists: a handful of points right around the robot (cells that carry an entry
    for almost every path) plus inside-robot points, for many headings."""
    rng = np.random.default_rng(91)
    xs = np.concatenate([rng.uniform(19.3, 20.7, 60), rng.uniform(0, 40, 40)]).astype(np.float32)
    ys = np.concatenate([rng.uniform(19.3, 20.7, 60), rng.uniform(0, 40, 40)]).astype(np.float32)
    poses = np.array([[20.0 + dx, 20.0 + dy, a] for a in np.linspace(-3.1, 3.1, 9) for dx in (-0.3, 0.0, 0.4)
                      for dy in (-0.2, 0.1)])
    got = _check(gpu_ctx, orc, oracle_ackermann, xs, ys, poses)
    assert (got == 0.0).any() and (got == 5.0).any()


def test_pinned_and_pageable_buffers_agree(gpu_ctx, capi):
    """The in-place path (pinned caller buffers, kernels read/write them over PCIe) and the staged-copy path (pageable
    buffers) return the same rows, for one-node, planner-sized and multi-sub-batch calls."""
    xs, ys = wl.uniform_cloud(30000, 60.0, seed=17)
    gpu_ctx.set_obstacles(xs, ys)
    for n in (1, 37, 1500, 5000):
        poses = wl.node_poses(n, 60.0, margin=1.0, seed=18 + n)
        pinned_in = capi.pinned_empty((n, 3), np.float64)
        pinned_in[:] = poses
        pinned_out = capi.pinned_empty((n, gpu_ctx.total_paths), np.float32)
        pinned_out[:] = -1.0
        a = gpu_ctx.eval_nodes(pinned_in, 5.0, out=pinned_out)
        b = gpu_ctx.eval_nodes(poses.copy(), 5.0)  # pageable numpy buffers
        assert a is pinned_out and np.array_equal(a, b)
        assert (a >= 0).all()


def test_large_batch_with_pinned_buffers(gpu_ctx, capi, orc, oracle_ackermann):
    """More than 1024 nodes with pinned buffers (the pipelined sub-batch path whose kernels store their rows straight
    into the caller's buffer): same values as the checker, also next to rows whose pose is not finite, repeatedly on
    the same buffers (rows of the previous call must not survive)."""
    xs, ys = wl.uniform_cloud(40000, 60.0, seed=27)
    gpu_ctx.set_obstacles(xs, ys)
    n = 3000
    pinned_in = capi.pinned_empty((n, 3), np.float64)
    pinned_out = capi.pinned_empty((n, gpu_ctx.total_paths), np.float32)
    for rep in range(3):
        poses = wl.node_poses(n, 60.0, margin=1.0, seed=28 + rep)
        poses[5, 2] = np.nan
        poses[77, 2] = np.inf
        poses[2999, 0] = np.nan
        pinned_in[:] = poses
        pinned_out[:] = -1.0
        got = gpu_ctx.eval_nodes(pinned_in, 5.0, out=pinned_out)
        assert got is pinned_out
        for i in (5, 77, 2999):  # no obstacle can be placed relative to such a node: nothing constrains its paths
            assert np.isinf(got[i]).all()
        sel = np.r_[0:40, 1000:1040, 2960:2999]
        want, _ = orc.eval_nodes(oracle_ackermann, xs, ys, poses[sel], 5.0, mode=1, nthreads=0)
        assert np.array_equal(free_dist(got[sel], oracle_ackermann), want)
        assert np.array_equal(got, gpu_ctx.eval_nodes(poses.copy(), 5.0))  # staged-copy path, pageable buffers


def test_full_size_parity_with_checker(gpu_ctx, orc, oracle_ackermann):
    """Config C3 at BASELINE's full size — 2^20 points, 4096 nodes, 2 x 121 paths = 991 232 edge evaluations — equals the
    checker's values bit for bit (the reference build needs ~2 s per 4096 nodes on 16 host threads)."""
    xs, ys = wl.uniform_cloud(1 << 20, 200.0)
    poses = wl.node_poses(4096, 200.0)
    got = _check(gpu_ctx, orc, oracle_ackermann, xs, ys, poses)
    assert got.shape == (4096, 242) and (got < 5.0).mean() > 0.5
    # the structured variant of SURVEY 8(d): 70 % of the points on wall segments at 0.05 m spacing
    xs, ys = wl.walls_cloud(1 << 20, 200.0)
    _check(gpu_ctx, orc, oracle_ackermann, xs, ys, poses[:1024])


@pytest.mark.parametrize("offset", [(1.0e5, -2.0e5), (-8191.7, 16384.3), (3.0e6, 3.0e6)])
def test_large_world_coordinates(gpu_ctx, orc, oracle_ackermann, offset):
    """Far from the origin the float cloud is coarse (ulp 0.008 m at 1e5 m, 0.25 m at 3e6 m) and the node position
    needs its hi/lo split in the FP32 filter: values must still be the reference's, bit for bit."""
    rng = np.random.default_rng(int(abs(offset[0])) % 1000)
    n = 6000
    xs = (offset[0] + rng.uniform(0, 30, n)).astype(np.float32)
    ys = (offset[1] + rng.uniform(0, 30, n)).astype(np.float32)
    poses = np.c_[offset[0] + rng.uniform(2, 28, 96), offset[1] + rng.uniform(2, 28, 96), rng.uniform(-np.pi, np.pi, 96)]
    _check(gpu_ctx, orc, oracle_ackermann, xs, ys, poses)


NEAR_CLOUDS = {
    "dense": lambda: wl.uniform_cloud(int(26.0 * 60 * 60), 60.0, seed=41),      # every path blocked nearby: one launch decides
    "medium": lambda: wl.uniform_cloud(int(6.0 * 60 * 60), 60.0, seed=42),      # some nodes decided, some not
    "sparse": lambda: wl.uniform_cloud(int(0.3 * 60 * 60), 60.0, seed=43),      # too few points nearby: straight to the list
    "walls": lambda: wl.walls_cloud(120000, 60.0, seed=44),
    "half_empty": lambda: tuple(a[: int(26.0 * 30 * 60)] * np.float32(s) for a, s in zip(wl.uniform_cloud(int(26.0 * 30 * 60), 60.0, seed=45), (0.5, 1.0))),
}


def _eval_resident(capi, ctx, poses, boxes, clip=5.0):
    """mppb_eval_nodes_boxed_device: inputs and rows resident in device memory (the entry point that takes the near-first form)."""
    import torch

    xycs = torch.from_numpy(capi.nodes_xycs_from_poses(np.ascontiguousarray(poses))).cuda()
    d_boxes = torch.from_numpy(boxes).cuda() if boxes is not None else None
    out = torch.full((len(poses), ctx.total_paths), -1.0, dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    rc = capi.lib().mppb_eval_nodes_boxed_device(ctx.h, len(poses), xycs.data_ptr(), d_boxes.data_ptr() if boxes is not None else None,
                                                  clip, out.data_ptr())
    assert rc == 0
    ctx.sync()
    return out.cpu().numpy()


@pytest.mark.parametrize("cloud", sorted(NEAR_CLOUDS))
@pytest.mark.parametrize("boxed", [False, True])
def test_near_first_evaluation_is_bit_identical(gpu_ctx, capi, orc, oracle_ackermann, cloud, boxed):
    """Batches of more than 256 nodes are evaluated near-first (two launches: the cloud bins around the node decide the
    rows whose paths are all blocked within the bound, a second launch takes the other nodes on the whole window):
    forced on (mode 2), adaptive (mode 1) and off (mode 0) return the same rows, equal to the checker's, on clouds
    where the first launch decides everything, something, and nothing."""
    xs, ys = NEAR_CLOUDS[cloud]()
    n = 700
    poses = wl.node_poses(n, 60.0, margin=-3.0, seed=46)  # some nodes outside the cloud's bounding box
    poses[3, 2] = np.nan
    rng = np.random.default_rng(47)
    boxes = None
    if boxed:
        boxes = np.stack([poses[:, 0] - rng.uniform(0.2, 6, n), poses[:, 1] - rng.uniform(0.2, 6, n),
                          poses[:, 0] + rng.uniform(0.2, 6, n), poses[:, 1] + rng.uniform(0.2, 6, n)], 1).astype(np.float32)
        boxes[3] = 0.0
    gpu_ctx.set_obstacles(xs, ys)
    rows = {}
    try:
        for mode in (0, 2, 1):
            gpu_ctx.set_grid_near(mode)
            before = gpu_ctx.grid_near_stats()
            rows[mode] = _eval_resident(capi, gpu_ctx, poses, boxes)
            after = gpu_ctx.grid_near_stats()
            if mode == 0:
                assert after[0] == before[0]
            elif mode == 2 and not os.environ.get("MPPB_GRID_STAGED"):  # (the staged form keeps to one launch)
                assert after[0] == before[0] + 1 and after[1] == before[1] + n
                pending = after[2] - before[2]
                assert 0 <= pending <= n
                if cloud == "dense" and not boxed:
                    assert pending < n // 3, pending  # (nodes at the cloud border see free space)
                if cloud == "sparse":
                    assert pending > n // 2, pending
    finally:
        gpu_ctx.set_grid_near(1)
    assert np.array_equal(rows[0].view(np.uint32), rows[2].view(np.uint32))
    assert np.array_equal(rows[0].view(np.uint32), rows[1].view(np.uint32))
    host = gpu_ctx.eval_nodes_boxed(poses, boxes, 5.0) if boxed else gpu_ctx.eval_nodes(poses, 5.0)  # host buffers: plain form
    assert np.array_equal(rows[0].view(np.uint32), host.view(np.uint32))
    if not boxed:
        sel = np.r_[0:48, n - 48:n]
        want, _ = orc.eval_nodes(oracle_ackermann, xs, ys, poses[sel], 5.0, mode=1, nthreads=0)
        assert np.array_equal(free_dist(rows[2][sel], oracle_ackermann), want)


def test_near_first_with_mixed_geometry_groups(capi, orc):
    """The near-first form with several kernel groups (grids of other extents and resolutions, other path counts): one
    near window and one bound — from the largest slack + cell diagonal over the groups — serve all of them."""
    from oracle.orc import ACKERMANN_SHAPE_X as X, ACKERMANN_SHAPE_Y as Y

    w = 60.0 * np.pi / 180.0
    specs = [(121, 5.0, 0.05, 1.0, w, +1.0), (31, 3.0, 0.1, 0.8, 1.3 * w, -1.0), (25, 2.35, 0.1, 0.6, 1.1 * w, +1.0)]
    optgs = [orc.diffdrive_c(n, rd, res, v, ww, K, X, Y) for (n, rd, res, v, ww, K) in specs]
    pptgs = [capi.PTG.diffdrive_c(n, rd, res, v, ww, K, X, Y) for (n, rd, res, v, ww, K) in specs]
    ctx = capi.Context(0)
    for p in pptgs:
        ctx.add_ptg(p)
    r, bound = None, None
    for density, seed in ((26.0, 91), (5.0, 92)):
        xs, ys = wl.uniform_cloud(int(density * 50 * 50), 50.0, seed=seed)
        poses = wl.node_poses(400, 50.0, margin=0.0, seed=seed + 10)
        ctx.set_obstacles(xs, ys)
        r, bound = ctx.grid_near_window(5.0)
        assert 0.0 < bound < r <= 2.0
        rows = {}
        for mode in (0, 2):
            ctx.set_grid_near(mode)
            s0 = ctx.grid_near_stats()
            rows[mode] = _eval_resident(capi, ctx, poses, None)
            s1 = ctx.grid_near_stats()
            if not os.environ.get("MPPB_GRID_STAGED"):  # (the staged form keeps to one launch)
                assert s1[0] - s0[0] == (1 if mode else 0)
        assert np.array_equal(rows[0].view(np.uint32), rows[2].view(np.uint32))
        sel = np.r_[0:40, 360:400]
        want, _ = orc.eval_nodes(optgs, xs, ys, poses[sel], 5.0, mode=1, nthreads=0)
        assert np.array_equal(free_dist(rows[2][sel], optgs), want)
    assert ctx.grid_near_window(3.0) == (0.0, 0.0)  # a clipping distance too small for a near window
    ctx.close()


def test_near_first_adaptive_switch(gpu_ctx, capi):
    """Mode 1 stops trying after a batch of which more than 2 % needed the second launch (a map with free space), starts
    afresh on a new cloud, and the rows do not depend on it."""
    xs, ys = wl.walls_cloud(20000, 60.0, seed=51)
    poses = wl.node_poses(600, 60.0, margin=1.0, seed=52)
    gpu_ctx.set_obstacles(xs, ys)
    gpu_ctx.set_grid_near(0)
    want = _eval_resident(capi, gpu_ctx, poses, None)
    gpu_ctx.set_grid_near(1)
    s0 = gpu_ctx.grid_near_stats()
    for _ in range(6):
        assert np.array_equal(_eval_resident(capi, gpu_ctx, poses, None).view(np.uint32), want.view(np.uint32))
    s1 = gpu_ctx.grid_near_stats()
    if os.environ.get("MPPB_GRID_STAGED"):
        return
    assert 1 <= s1[0] - s0[0] <= 2  # the first call (and at most the one issued before its count was in) took two launches
    # a dense cloud afterwards: the switch starts afresh and stays on
    xs, ys = wl.uniform_cloud(int(26.0 * 60 * 60), 60.0, seed=53)
    poses = wl.node_poses(600, 60.0, margin=6.0, seed=54)
    gpu_ctx.set_obstacles(xs, ys)
    for _ in range(4):
        _eval_resident(capi, gpu_ctx, poses, None)
    s2 = gpu_ctx.grid_near_stats()
    assert s2[0] - s1[0] == 4 and s2[2] - s1[2] <= 4 * 12


def test_many_nodes_in_one_batch(gpu_ctx):
    """A batch far above the sub-batch size and above 65 535 nodes: every row equals the row of its pose evaluated
    alone in a small batch."""
    xs, ys = wl.uniform_cloud(40000, 60.0, seed=51)
    gpu_ctx.set_obstacles(xs, ys)
    base = wl.node_poses(257, 60.0, margin=1.0, seed=52)
    rows = gpu_ctx.eval_nodes(base, 5.0)
    reps = 70000 // len(base) + 1
    big = np.tile(base, (reps, 1))
    out = gpu_ctx.eval_nodes(big, 5.0)
    assert out.shape[0] > 65535
    assert np.array_equal(out, np.tile(rows, (reps, 1)))


def test_more_than_128_points_inside_the_robot_shape(gpu_ctx, orc, oracle_ackermann):
    """The kernel queues obstacle points that fall inside the robot polygon (COLL_BEH_BACK_AWAY) in a 128-entry list per
    node; beyond that it applies them on the spot. 600 points inside the shape of each node, plus a uniform background
    (dense window form) and alone (sparse window form, <= 256 points is exceeded on purpose by the 600)."""
    rng = np.random.default_rng(31)
    extent = 30.0
    poses = wl.node_poses(12, extent, margin=3.0, seed=32)
    px, py = [], []
    for x, y, phi in poses:
        # a disc of radius 0.09 around the local point (0.05, 0) lies inside the polygon (x in [-0.1, 0.3], |y| <= 0.1)
        r, a = 0.09 * np.sqrt(rng.uniform(size=600)), rng.uniform(-np.pi, np.pi, 600)
        lx, ly = 0.05 + r * np.cos(a), r * np.sin(a)
        px.append(x + lx * np.cos(phi) - ly * np.sin(phi))
        py.append(y + lx * np.sin(phi) + ly * np.cos(phi))
    inx, iny = np.concatenate(px).astype(np.float32), np.concatenate(py).astype(np.float32)
    inside_frac = np.mean([oracle_ackermann[0].inside_shape(0.05 + 0.08 * np.cos(t), 0.08 * np.sin(t)) for t in np.linspace(0, 6.2, 40)])
    assert inside_frac == 1.0
    bx, by = wl.uniform_cloud(int(2.0 * extent * extent), extent, seed=33)
    for xs, ys in ((inx, iny), (np.concatenate([inx, bx]), np.concatenate([iny, by]))):
        # (near the origin every path's cell distance is below the robot radius, which BACK_AWAY ignores: the 600 points
        # leave most paths free — what is tested is that 600 queued-or-overflowed points give the reference's rows)
        _check(gpu_ctx, orc, oracle_ackermann, xs, ys, poses)


def test_more_than_128_long_cell_lists_in_a_sparse_window(gpu_ctx, capi, orc, oracle_ackermann, product_ackermann):
    """Sparse windows (<= 256 points) scatter the by-cell lists of the occupied cells; lists longer than 32 entries
    (cells near the robot carry an entry for most paths) are queued, 128 per node, and folded by the whole block;
    beyond 128 the thread folds them itself. ~230 points in distinct cells with long lists around each node."""
    g0, g1 = product_ackermann[0].grid(), product_ackermann[1].grid()
    lens = (np.diff(g0["offsets"]) + np.diff(g1["offsets"])).reshape(g0["ny"], g0["nx"])  # merged list length per cell
    cy, cx = np.nonzero(lens > 32)
    # cell centres outside the robot's bounding circle (so that the points are plain obstacles, not inside-shape ones)
    ccx = g0["xmin"] + (cx + 0.5) * g0["res"]
    ccy = g0["ymin"] + (cy + 0.5) * g0["res"]
    keep = np.hypot(ccx, ccy) > g0["max_radius"] + 0.08
    ccx, ccy = ccx[keep], ccy[keep]
    assert len(ccx) > 400
    rng = np.random.default_rng(41)
    poses = wl.node_poses(6, 300.0, margin=20.0, seed=42)  # far apart: each window holds its own points only
    px, py = [], []
    for x, y, phi in poses:
        pick = rng.choice(len(ccx), size=230, replace=False)
        lx, ly = ccx[pick], ccy[pick]
        px.append(x + lx * np.cos(phi) - ly * np.sin(phi))
        py.append(y + lx * np.sin(phi) + ly * np.cos(phi))
    xs, ys = np.concatenate(px).astype(np.float32), np.concatenate(py).astype(np.float32)
    got = _check(gpu_ctx, orc, oracle_ackermann, xs, ys, poses)
    assert (got < 5.0).mean() > 0.5


def test_collision_grid_larger_than_the_shared_memory_bitmap(capi, orc):
    """A 35 m reference distance at 0.05 m cells is a 1400 x 1400 grid: its occupancy bitmap (245 KB) does not fit
    shared memory, so the kernel takes its form without a bitmap (every point folds its cell's list at once). Same rows
    as the checker, whose tables are built independently."""
    p_ptgs = capi.ackermann_ptgs(num_paths=11, ref_distance=35.0)
    o_ptgs = orc.ackermann_ptgs(num_paths=11, ref_distance=35.0)
    for p, o in zip(p_ptgs, o_ptgs):
        gp, go = p.grid(), o.grid()
        assert gp["nx"] == 1400 and np.array_equal(gp["offsets"], go["offsets"]) and np.array_equal(gp["ds"], go["ds"])
    ctx = capi.Context(0)
    for p in p_ptgs:
        ctx.add_ptg(p)
    extent = 160.0
    xs, ys = wl.uniform_cloud(int(0.4 * extent * extent), extent, seed=51)
    poses = wl.node_poses(24, extent, margin=5.0, seed=52)
    got = _check(ctx, orc, o_ptgs, xs, ys, poses, clip=35.0)
    assert (got < 35.0).any()
    # per-node world boxes on the same form
    boxes = np.stack([poses[:, 0] - 20, poses[:, 1] - 20, poses[:, 0] + 20, poses[:, 1] + 20], 1).astype(np.float32)
    out = np.empty((len(poses), ctx.total_paths), dtype=np.float32)
    assert capi.lib().mppb_eval_nodes_boxed(ctx.h, len(poses), np.ascontiguousarray(poses).ctypes.data, boxes.ctypes.data, 35.0,
                                             out.ctypes.data) == 0
    for i in range(0, len(poses), 6):
        m = (xs >= boxes[i, 0]) & (xs <= boxes[i, 2]) & (ys >= boxes[i, 1]) & (ys <= boxes[i, 3])
        ref, _ = orc.eval_nodes(o_ptgs, xs[m], ys[m], poses[i:i + 1], 35.0, mode=1, nthreads=0)
        assert np.array_equal(free_dist(out[i:i + 1], o_ptgs), ref)
    ctx.close()


def test_staged_form_of_the_kernel_is_bit_identical():
    """tpobs_grid_kernel's STAGED form (points through cp.async.bulk + mbarrier into a shared-memory ring; selected
    with MPPB_GRID_STAGED=1, read once per process) gives the same rows as the default form: the parity tests of this
    file and the golden vectors, in a child process with the variable set."""
    import subprocess
    import sys

    here = os.path.dirname(os.path.abspath(__file__))
    env = dict(os.environ, MPPB_GRID_STAGED="1")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(here, "test_tpobs_grid_gpu.py"),
                        os.path.join(here, "test_golden_gpu.py"), "-x", "-q", "-k", "not staged_form"], env=env,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:]
