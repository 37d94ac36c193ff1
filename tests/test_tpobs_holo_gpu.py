"""GPU parity of stage (a)+(b) for the HolonomicBlend PTG: libmpp_b200 (through the C ABI) vs the CPU
oracle on the same seeded candidates (node pose + velocity, trajectory k, trimmed speed).

Bar: the arithmetic is FP64 with the reference's operation order; only acos/cos inside the cubic
solver come from a different libm, so distances must agree to 1e-9 relative (north_star allows 1e-4)
and the set of unconstrained candidates (+inf) and of blocked-at-zero candidates must be identical."""
import os

import numpy as np
import pytest

from mrpt_path_planning_b200 import capi as _capi, workloads as wl

pytestmark = pytest.mark.gpu
RTOL = 1e-9


@pytest.fixture(scope="module")
def holo_ctx(capi):
    ptg = capi.holonomic_ptgs()[0]
    ctx = capi.Context(0)
    ctx.add_ptg(ptg)
    yield ctx, ptg
    ctx.close()


def make_candidates(ptg, nodes, rel_targets, ks, speeds):
    """Candidate list in find_feasible order (speed-major), parameters from the product's host PTG."""
    cn, ck, cs, rows = [], [], [], []
    for b, nd in enumerate(nodes):
        c, s = np.cos(-nd[2]), np.sin(-nd[2])
        vloc = [nd[3] * c - nd[4] * s, nd[3] * s + nd[4] * c, nd[5]]
        for sp in speeds:
            ptg.set_trim_speed(sp)
            ptg.update_dyn_state(vloc, rel_targets[b], 0.0, force=True)
            for k in ks[b]:
                p = ptg.holo_params(int(k))
                rows.append((b, 0, p.T_ramp, p.vxi, p.vyi, p.vxf, p.vyf, p.vf))
                cn.append(b)
                ck.append(int(k))
                cs.append(sp)
    return np.array(rows, dtype=_capi.HOLO_CAND_DTYPE), cn, ck, cs


def check(holo_ctx, orc, xs, ys, nodes, rel_targets, ks, speeds=(1 / 3, 2 / 3, 1.0), clip=5.0):
    ctx, ptg = holo_ctx
    optg = orc.holonomic_ptgs()[0]
    cands, cn, ck, cs = make_candidates(ptg, nodes, rel_targets, ks, speeds)
    ctx.set_obstacles(xs, ys)
    got = ctx.eval_holo_candidates(nodes[:, :3], cands, clip)
    ref, _, _ = orc.eval_candidates(optg, xs, ys, nodes, rel_targets, cn, ck, cs, clip, want_free=False)
    assert np.array_equal(np.isinf(got), np.isinf(ref))
    assert np.array_equal(got == 0.0, ref == 0.0)
    fin = np.isfinite(ref)
    err = np.abs(got[fin] - ref[fin]) / np.maximum(np.abs(ref[fin]), 1e-12)
    assert err.size == 0 or err.max() <= RTOL, "max rel err %g" % err.max()
    return got, ref


def random_nodes(rng, b, extent, moving=True):
    nodes = np.zeros((b, 6))
    nodes[:, 0] = rng.uniform(1, extent - 1, b)
    nodes[:, 1] = rng.uniform(1, extent - 1, b)
    nodes[:, 2] = rng.uniform(-np.pi, np.pi, b)
    if moving:
        sp = rng.uniform(0, 1.0, b) * (rng.uniform(size=b) < 0.7)
        th = rng.uniform(-np.pi, np.pi, b)
        nodes[:, 3], nodes[:, 4] = sp * np.cos(th), sp * np.sin(th)
    rel = np.stack([rng.uniform(-8, 8, b), rng.uniform(-8, 8, b), rng.uniform(-3, 3, b)], axis=1)
    return nodes, rel


def test_c1_cloud_parity(holo_ctx, orc):
    """the reference's own demo cloud (share/obstacles_01.txt, 155 points), planner k-subset"""
    obs = np.loadtxt(os.path.join(os.path.dirname(__file__), "golden", "obstacles_01.txt")).astype(np.float32)
    rng = np.random.default_rng(5)
    nodes, rel = random_nodes(rng, 48, 3.0)
    nodes[:, 0] = rng.uniform(-1.0, 5.5, 48)
    nodes[:, 1] = rng.uniform(-1.5, 4.0, 48)
    k13 = [0, 15, 31, 47, 63, 79, 95, 110, 126, 142, 158, 174, 190]
    got, ref = check(holo_ctx, orc, obs[:, 0].copy(), obs[:, 1].copy(), nodes, rel, [k13] * 48)
    assert np.isfinite(ref).any() and np.isinf(ref).any()


@pytest.mark.parametrize("density,seed", [(0.2, 1), (2.0, 2), (26.0, 3)])
def test_uniform_cloud_all_paths(holo_ctx, orc, density, seed):
    extent = 30.0
    xs, ys = wl.uniform_cloud(int(density * extent * extent), extent, seed=seed)
    rng = np.random.default_rng(seed + 50)
    nodes, rel = random_nodes(rng, 6, extent)
    check(holo_ctx, orc, xs, ys, nodes, rel, [list(range(0, 191, 3))] * 6)


def test_zero_velocity_straight_known_answer(holo_ctx):
    """single obstacle dead ahead of a node at rest: free distance = ox - R in ramp and cruise phase"""
    ctx, ptg = holo_ctx
    nodes = np.array([[10.0, 10.0, 0.0, 0, 0, 0]])
    for ox in (0.3, 0.45, 2.0, 4.5):
        ctx.set_obstacles(np.float32([10.0 + ox]), np.float32([10.0]))
        cands, *_ = make_candidates(ptg, nodes, [[20.0, 0, 0]], [[95]], [1.0])
        got = ctx.eval_holo_candidates(nodes[:, :3], cands, 5.0)
        assert abs(got[0] - (ox - 0.15)) < 1e-6  # the float narrowing of the local point costs ~1e-7


def test_empty_cloud_and_empty_batch(holo_ctx):
    ctx, ptg = holo_ctx
    ctx.set_obstacles(np.zeros(0, np.float32), np.zeros(0, np.float32))
    nodes = np.array([[1.0, 2.0, 0.3, 0, 0, 0]])
    cands, *_ = make_candidates(ptg, nodes, [[3.0, 1.0, 0]], [[0, 95, 190]], [1.0])
    assert np.isinf(ctx.eval_holo_candidates(nodes[:, :3], cands, 5.0)).all()
    assert ctx.eval_holo_candidates(nodes[:, :3], cands[:0], 5.0).shape == (0,)


def test_bad_candidate_is_rejected(holo_ctx, capi):
    ctx, ptg = holo_ctx
    nodes = np.array([[1.0, 2.0, 0.3, 0, 0, 0]])
    cands, *_ = make_candidates(ptg, nodes, [[3.0, 1.0, 0]], [[5]], [1.0])
    bad = cands.copy()
    bad["node"] = 7
    with pytest.raises(capi.MppbError):
        ctx.eval_holo_candidates(nodes[:, :3], bad, 5.0)
    bad = cands.copy()
    bad["ptg_index"] = 3
    with pytest.raises(capi.MppbError):
        ctx.eval_holo_candidates(nodes[:, :3], bad, 5.0)


def test_blocked_decisions_match(holo_ctx, orc):
    """the free/blocked decision of TPS_Astar.cpp:749 for the planner's sample timestamps"""
    ctx, ptg = holo_ctx
    optg = orc.holonomic_ptgs()[0]
    extent = 20.0
    xs, ys = wl.uniform_cloud(400, extent, seed=9)
    rng = np.random.default_rng(77)
    nodes, rel = random_nodes(rng, 24, extent)
    ks = [[0, 15, 31, 47, 63, 79, 95, 110, 126, 142, 158, 174, 190]] * 24
    got, ref = check(holo_ctx, orc, xs, ys, nodes, rel, ks)
    cands, cn, ck, cs = make_candidates(ptg, nodes, rel, ks, (1 / 3, 2 / 3, 1.0))
    n_dec = 0
    for i in range(len(cn)):
        b = cn[i]
        c, s = np.cos(-nodes[b, 2]), np.sin(-nodes[b, 2])
        vloc = [nodes[b, 3] * c - nodes[b, 4] * s, nodes[b, 3] * s + nodes[b, 4] * c, nodes[b, 5]]
        for P in (ptg, optg):
            P.set_trim_speed(cs[i])
            P.update_dyn_state(vloc, rel[b], 0.0, force=True)
        init_g, init_r = ptg.init_dist(ck[i]), optg.init_dist(ck[i])
        assert init_g == init_r
        for step in (50, 150, 400):
            if step >= ptg.step_count(ck[i]):
                continue
            d = ptg.path_pose(ck[i], step)[3]
            assert (d >= min(init_g, got[i])) == (d >= min(init_r, ref[i]))
            n_dec += 1
    assert n_dec > 1000


@pytest.mark.parametrize("density,seed", [(1.0, 21), (8.0, 22), (40.0, 23)])
def test_filter_is_conservative(holo_ctx, density, seed):
    """The geometric pre-filter only drops points that cannot change the result: the kernel with the filter and the
    kernel without it (every in-window point through the exact FP64 chain) return the same bits, for every path of the
    PTG, moving and resting nodes, all three trimmed speeds."""
    ctx, ptg = holo_ctx
    extent = 24.0
    xs, ys = wl.uniform_cloud(int(density * extent * extent), extent, seed=seed)
    rng = np.random.default_rng(seed + 7)
    nodes, rel = random_nodes(rng, 10, extent)
    nodes[::3, 3:] = 0.0  # nodes at rest: degenerate control triangle
    nodes[1, 3:5] = [1.0, 0.0]
    cands, *_ = make_candidates(ptg, nodes, rel, [list(range(191))] * len(nodes), (1 / 3, 2 / 3, 1.0))
    ctx.set_obstacles(xs, ys)
    try:
        ctx.set_holo_filter(False)
        plain = ctx.eval_holo_candidates(nodes[:, :3], cands, 5.0)
    finally:
        ctx.set_holo_filter(True)
    filt = ctx.eval_holo_candidates(nodes[:, :3], cands, 5.0)
    assert np.array_equal(plain.view(np.uint64), filt.view(np.uint64))
    assert np.isfinite(plain).any()


@pytest.mark.parametrize("offset,extent,clip", [(0.0, 24.0, 5.0), (-3000.0, 30.0, 7.5), (1.0e5, 40.0, 5.0), (0.0, 6.0, 12.0)])
def test_bin_culling_is_conservative(holo_ctx, offset, extent, clip):
    """Bin culling (the warp walks only the cloud bins its candidate's inflated hull can reach) returns the same bits
    as the pre-filter over the whole window (mode 2) and as no filter at all (mode 0): nodes inside, at the border of
    and outside the cloud's bin grid (first/last bin rows and columns catch everything beyond), clouds far from the
    origin (float coordinates with a coarse ulp), windows larger than the cloud."""
    ctx, ptg = holo_ctx
    xs, ys = wl.uniform_cloud(int(12.0 * extent * extent), extent, seed=int(extent))
    xs = (xs + np.float32(offset)).astype(np.float32)
    ys = (ys - np.float32(offset)).astype(np.float32)
    rng = np.random.default_rng(int(extent) + 1)
    nodes, rel = random_nodes(rng, 16, extent)
    nodes[:4, 0] = [-1.0, extent + 1.0, 0.01, extent - 0.01]  # outside and on the border of the bin grid
    nodes[4:8, 1] = [-2.0, extent + 0.5, 0.0, extent]
    nodes[::5, 3:] = 0.0
    nodes[:, 0] += offset
    nodes[:, 1] -= offset
    cands, *_ = make_candidates(ptg, nodes, rel, [list(range(0, 191, 3))] * len(nodes), (1 / 3, 1.0))
    ctx.set_obstacles(xs, ys)
    res = {}
    try:
        for mode in (0, 2, 1):
            ctx.set_holo_filter(mode)
            res[mode] = ctx.eval_holo_candidates(nodes[:, :3], cands, clip)
    finally:
        ctx.set_holo_filter(1)
    assert np.array_equal(res[0].view(np.uint64), res[2].view(np.uint64))
    assert np.array_equal(res[0].view(np.uint64), res[1].view(np.uint64))
    assert np.isfinite(res[0]).any() and np.isinf(res[0]).any()


HOLO_VARIANTS = [
    # num_paths, ref_distance, T_ramp_max, v_max, w_max, radius, expr_V, expr_W
    (61, 4.0, 0.5, 0.8, np.deg2rad(90.0), 0.30, "", ""),
    (101, 6.0, 2.0, 1.5, np.deg2rad(30.0), 0.10, "V_MAX * max(0.2, trimmable_speed) * (1 - 0.3*abs(dir)/PI)",
     "W_MAX * min(1.0, abs(dir))"),
]


@pytest.mark.parametrize("variant", range(len(HOLO_VARIANTS)))
def test_other_ptg_parameters(capi, orc, variant):
    """HolonomicBlend with other parameters and user expressions than the demo's (ramp time, speeds, radius, path
    count, clip = its own refDistance): kernel (with its pre-filter) vs the checker, and filter on == filter off."""
    v = HOLO_VARIANTS[variant]
    ptg = capi.PTG.holonomic_blend(*v)
    optg = orc.holonomic_blend(*v)
    ctx = capi.Context(0)
    ctx.add_ptg(ptg)
    rng = np.random.default_rng(60 + variant)
    xs, ys = wl.uniform_cloud(4000, 25.0, seed=61 + variant)
    nodes, rel = random_nodes(rng, 8, 25.0)
    ks = [list(range(0, v[0], 2))] * len(nodes)
    cands, cn, ck, cs = make_candidates(ptg, nodes, rel, ks, (0.5, 1.0))
    ctx.set_obstacles(xs, ys)
    got = ctx.eval_holo_candidates(nodes[:, :3], cands, v[1])
    ctx.set_holo_filter(False)
    plain = ctx.eval_holo_candidates(nodes[:, :3], cands, v[1])
    ctx.close()
    assert np.array_equal(got.view(np.uint64), plain.view(np.uint64))
    ref, _, _ = orc.eval_candidates(optg, xs, ys, nodes, rel, cn, ck, cs, v[1], want_free=False)
    assert np.array_equal(np.isinf(got), np.isinf(ref)) and np.array_equal(got == 0.0, ref == 0.0)
    fin = np.isfinite(ref)
    assert fin.any() and (np.abs(got[fin] - ref[fin]) / np.maximum(np.abs(ref[fin]), 1e-12)).max() <= RTOL


@pytest.mark.parametrize("n_nodes,boxed", [(1, False), (1, True), (7, True), (300, False)])
def test_split_calls_equal_blocking_calls(capi, product_ackermann, n_nodes, boxed):
    """mppb_eval_nodes_boxed_begin + mppb_eval_holo_candidates_boxed_begin in flight together, one mppb_sync: the
    same bits as the two blocking calls (one context holding both PTG kinds, as a mixed planner would)."""
    ptg = capi.holonomic_ptgs()[0]
    ctx = capi.Context(0)
    for p in product_ackermann:
        ctx.add_ptg(p)
    holo_index = ctx.num_ptgs
    ctx.add_ptg(ptg)
    rng = np.random.default_rng(100 + n_nodes)
    xs, ys = wl.uniform_cloud(40000, 40.0, seed=31)
    ctx.set_obstacles(xs, ys)
    nodes, rel = random_nodes(rng, n_nodes, 40.0)
    ks = [rng.choice(ptg.num_paths, 9, replace=False) for _ in range(n_nodes)]
    cands, _, _, _ = make_candidates(ptg, nodes, rel, ks, (1 / 3, 1.0))
    cands["ptg_index"] = holo_index
    poses = nodes[:, :3]
    boxes = None
    if boxed:
        boxes = np.stack([poses[:, 0] - 3.0, poses[:, 1] - 2.5, poses[:, 0] + 4.0, poses[:, 1] + 3.5], 1).astype(np.float32)
        want_rows = ctx.eval_nodes_boxed(poses, boxes, 5.0)
        want_holo = ctx.eval_holo_candidates_boxed(poses, boxes, cands, 5.0)
    else:
        want_rows = ctx.eval_nodes(poses, 5.0)
        want_holo = ctx.eval_holo_candidates(poses, cands, 5.0)
    assert np.isfinite(want_holo).any() and np.isfinite(want_rows).any()
    for _ in range(3):  # back to back: the staging buffers are reused
        rows, hout = ctx.eval_both_split(poses, boxes, cands, 5.0)
        assert np.array_equal(rows, want_rows)
        assert np.array_equal(hout, want_holo)
    ctx.close()


@pytest.mark.parametrize("density,boxed,filt", [(3.0, False, True), (26.0, True, True), (90.0, False, True), (8.0, True, False)])
def test_cta_per_node_kernel_equals_warp_per_candidate_kernel(holo_ctx, capi, orc, density, boxed, filt):
    """mppb_set_holo_by_node: batches whose candidates are grouped by node run tpobs_holo_node_kernel (one CTA per node:
    the window is walked once for all the node's candidates) instead of the warp-per-candidate kernel. Same points, same
    per-point arithmetic: the same bits. 90 points/m^2 puts > 6144 points into a window (several passes)."""
    ctx, ptg = holo_ctx
    extent = 26.0
    xs, ys = wl.uniform_cloud(int(density * extent * extent), extent, seed=int(density) + 3)
    rng = np.random.default_rng(int(density) + 9)
    n = 72
    nodes, rel = random_nodes(rng, n, extent)
    ks = [sorted(rng.choice(191, size=int(rng.integers(1, 20)), replace=False).tolist()) for _ in range(n)]
    ks[5] = []  # a node without candidates
    cands, cn, ck, cs = make_candidates(ptg, nodes, rel, ks, (1 / 3, 1.0))
    boxes = None
    if boxed:
        boxes = np.stack([nodes[:, 0] - rng.uniform(1, 6, n), nodes[:, 1] - rng.uniform(1, 6, n),
                          nodes[:, 0] + rng.uniform(1, 6, n), nodes[:, 1] + rng.uniform(1, 6, n)], 1).astype(np.float32)
    ctx.set_obstacles(xs, ys)
    L = capi.lib()

    def run(node_sel):
        sel = np.isin(cands["node"], node_sel)
        sub = cands[sel].copy()
        remap = {int(b): i for i, b in enumerate(node_sel)}
        sub["node"] = [remap[int(b)] for b in sub["node"]]
        poses = np.ascontiguousarray(nodes[node_sel, :3])
        bx = np.ascontiguousarray(boxes[node_sel]) if boxes is not None else None
        out = np.empty(len(sub), dtype=np.float64)
        rc = L.mppb_eval_holo_candidates_boxed(ctx.h, len(poses), poses.ctypes.data, bx.ctypes.data if bx is not None else None,
                                               len(sub), sub.ctypes.data, 5.0, out.ctypes.data)
        assert rc == 0
        return sel, out

    try:
        ctx.set_holo_filter(filt)
        ctx.set_holo_by_node(True)
        _, whole = run(list(range(n)))                       # CTA per node
        ctx.set_holo_by_node(False)
        parts = np.empty_like(whole)
        for lo in range(0, n, 24):                            # warp per candidate, 24 nodes at a time
            sel, o = run(list(range(lo, lo + 24)))
            parts[sel] = o
    finally:
        ctx.set_holo_filter(True)
        ctx.set_holo_by_node(False)
    assert np.array_equal(whole.view(np.uint64), parts.view(np.uint64))
    assert np.isfinite(whole).any()
    if not boxed:
        ref, _, _ = orc.eval_candidates(orc.holonomic_ptgs()[0], xs, ys, nodes, rel, cn, ck, cs, 5.0, want_free=False)
        assert np.array_equal(np.isinf(whole), np.isinf(ref)) and np.array_equal(whole == 0.0, ref == 0.0)
        fin = np.isfinite(ref)
        assert (np.abs(whole[fin] - ref[fin]) / np.maximum(np.abs(ref[fin]), 1e-12)).max() <= RTOL
