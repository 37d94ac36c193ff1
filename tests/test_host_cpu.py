"""CPU tests of the product's host-side mirror (no GPU, no compute calls into CUDA): the HolonomicBlend host PTG
equals the oracle bit for bit, the config readers parse the reference's own demo files, expressions evaluate."""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SHARE = os.path.join(HERE, "golden", "share")


def test_holonomic_host_ptg_equals_oracle(capi, orc):
    h = capi.holonomic_ptgs()[0]
    o = orc.holonomic_ptgs()[0]
    rng = np.random.default_rng(0)
    for _ in range(300):
        vel = rng.uniform(-1, 1, 3) * rng.choice([0, 1])
        tgt = [rng.uniform(-5, 5), rng.uniform(-5, 5), rng.uniform(-3, 3)]
        sp = rng.choice([1 / 3, 2 / 3, 1.0])
        for p in (h, o):
            p.set_trim_speed(sp)
            p.update_dyn_state(vel, tgt, 0.0, force=True)
        k, st = int(rng.integers(0, 191)), int(rng.integers(1, 400))
        assert np.array_equal(h.path_pose(k, st), o.path_pose(k, st))
        assert h.step_count(k) == o.step_count(k)
        assert h.step_for_dist(k, 1.7) == o.step_for_dist(k, 1.7)
        assert h.inverse_map(tgt[0], tgt[1]) == o.inverse_map(tgt[0], tgt[1])
        assert np.array_equal(h.path_twist(k, st), o.path_twist(k, st))
        assert h.init_dist(k) == o.init_dist(k)


def test_holo_step_count_cache_survives_speed_change(capi, orc):
    """the reference does not invalidate the step-count cache when the trimmed speed changes"""
    for mod in (capi, orc):
        p = mod.holonomic_ptgs()[0]
        p.set_trim_speed(1.0)
        p.update_dyn_state((0, 0, 0), (3.0, 1.0, 0.0), 0.0, force=True)
        n_full = p.step_count(40)
        p.set_trim_speed(1 / 3)
        assert p.step_count(40) == n_full  # cached
        p.update_dyn_state((0.1, 0, 0), (3.0, 1.0, 0.0), 0.0)  # new dynamic state: cache dropped
        assert p.step_count(40) != n_full


def test_holo_candidate_params(capi):
    p = capi.holonomic_ptgs()[0]
    p.set_trim_speed(2 / 3)
    p.update_dyn_state((0.2, -0.1, 0.0), (2.0, 0.0, 0.0), 0.0, force=True)
    c = p.holo_params(95)  # dir = 0
    assert c.T_ramp == 1.0 and c.vxi == 0.2 and c.vyi == -0.1
    assert abs(c.vf - 2 / 3) < 1e-15 and abs(c.vxf - 2 / 3) < 1e-15 and c.vyf == 0.0


def test_cli_binary_is_built(capi):
    from mrpt_path_planning_b200 import build

    capi.lib()
    assert os.access(build.CLI_PATH, os.X_OK)


def test_worker_pool_runs_every_item_once(capi):
    """the pool behind plan_batch (slices with affinity + helping): every index exactly once per round, for counts
    below, at and far above the worker count; an exception in a work item reaches the caller and the pool survives"""
    for n_threads, count in ((1, 5), (4, 1), (4, 3), (4, 4), (7, 50), (16, 512), (3, 1000)):
        assert capi.lib().mppb_selftest_worker_pool(n_threads, count, 20) == 0, (n_threads, count)


def test_holo_direction_memo_is_transparent(capi, orc):
    """HolonomicBlend keeps its per-direction parameters in a small direct-mapped memo per dynamic state: any order of
    (k, trimmed speed) queries — slot collisions included — gives the oracle's values"""
    h = capi.holonomic_ptgs()[0]
    o = orc.holonomic_ptgs()[0]
    rng = np.random.default_rng(5)
    for state in range(4):
        vel = rng.uniform(-0.8, 0.8, 3)
        tgt = [rng.uniform(-5, 5), rng.uniform(-5, 5), rng.uniform(-3, 3)]
        for p in (h, o):
            p.update_dyn_state(vel, tgt, 0.0, force=True)
        for _ in range(1500):
            sp = float(rng.choice([1 / 3, 2 / 3, 1.0, 0.37, 0.81]))
            k, st = int(rng.integers(0, 191)), int(rng.integers(1, 300))
            for p in (h, o):
                p.set_trim_speed(sp)
            assert np.array_equal(h.path_pose(k, st), o.path_pose(k, st))
            assert h.init_dist(k) == o.init_dist(k)
        # a fresh PTG (empty memo) agrees on the candidate records
        f = capi.holonomic_ptgs()[0]
        f.update_dyn_state(vel, tgt, 0.0, force=True)
        for k in range(0, 191, 7):
            for sp in (1 / 3, 1.0):
                h.set_trim_speed(sp)
                f.set_trim_speed(sp)
                a, b = h.holo_params(k), f.holo_params(k)
                assert (a.T_ramp, a.vxi, a.vyi, a.vxf, a.vyf, a.vf) == (b.T_ramp, b.vxi, b.vyi, b.vxf, b.vyf, b.vf)


def test_holo_memo_with_state_reading_expressions(capi, orc):
    """|V| and |omega| formulas that read the dynamic state (target, current velocity): memoised parameters must be
    dropped at every state change; formulas of dir and the trimmed speed only keep theirs. Either way the values
    are the oracle's over a sequence of state changes with repeated (k, speed) queries."""
    variants = [
        (61, 4.0, 0.5, 0.8, np.deg2rad(90.0), 0.30, "V_MAX * trimmable_speed * min(1.0, 0.3 + target_dist/4)",
         "W_MAX * min(1.0, abs(dir - target_dir) + 0.2*abs(vxi) + 0.1*target_rel_speed)"),
        (61, 4.0, 0.5, 0.8, np.deg2rad(90.0), 0.30, "V_MAX * max(0.2, trimmable_speed) * (1 - 0.3*abs(dir)/PI)",
         "W_MAX * min(1.0, abs(dir))"),
    ]
    rng = np.random.default_rng(11)
    for v in variants:
        h, o = capi.PTG.holonomic_blend(*v), orc.holonomic_blend(*v)
        queries = [(int(rng.integers(0, 61)), float(rng.choice([0.5, 1.0])), int(rng.integers(1, 200))) for _ in range(40)]
        for state in range(12):
            vel = rng.uniform(-0.6, 0.6, 3)
            tgt = [rng.uniform(-4, 4), rng.uniform(-4, 4), rng.uniform(-3, 3)]
            trs = float(rng.uniform(0, 1))
            for p in (h, o):
                p.update_dyn_state(vel, tgt, trs)
            for k, sp, st in queries + queries[::-1]:
                for p in (h, o):
                    p.set_trim_speed(sp)
                assert np.array_equal(h.path_pose(k, st), o.path_pose(k, st)), (v[6], state, k, sp, st)
                assert h.init_dist(k) == o.init_dist(k)
                assert h.step_for_dist(k, 1.1) == o.step_for_dist(k, 1.1)


def test_open_set_pops_in_multimap_order(capi):
    """the planner's heap open set leaves in the order of the reference's std::multimap (smallest key first, equal
    keys in insertion order), whatever the mix of insertions and pops"""
    for seed, ops, keys in ((1, 10, 1), (2, 1000, 3), (3, 20000, 50), (4, 20000, 100000), (5, 5000, 2)):
        assert capi.lib().mppb_selftest_open_set(seed, ops, keys) == 0, (seed, ops, keys)


def test_lattice_index_is_a_stable_map(capi):
    """the planner's open-addressing lattice index behaves like the reference's unordered_map<NodeCoords, Node> under
    indexing: one node per coordinate, stable addresses through table growth"""
    for seed, ops, rng_ in ((1, 100, 2), (2, 50000, 40), (3, 200000, 400), (4, 3000, 100000)):
        assert capi.lib().mppb_selftest_lattice_index(seed, ops, rng_) == 0, (seed, ops, rng_)


def test_cell_order_equals_unordered_map_iteration_order(capi):
    """on the device-expansion path the host replays the ORDER in which the reference's unordered_map of reached cells
    is iterated (it assigns the node ids) instead of filling a real container: same order for any key set, through
    every rehash (14th, 30th, 60th, 128th ... key)"""
    for seed, trials, max_keys in ((1, 2000, 13), (2, 2000, 14), (3, 3000, 31), (4, 3000, 61), (5, 3000, 90),
                                   (6, 1000, 300), (7, 200, 2000)):
        assert capi.lib().mppb_selftest_cell_order(seed, trials, max_keys) == 0, (seed, trials, max_keys)


def test_sincos_helper_thread_is_transparent(capi):
    """large host batches share the heading cos/sin between the calling thread and a helper thread (spinning for 2 ms
    after a request, asleep afterwards): every row equals the single-threaded one bit for bit, at any split point,
    with small and huge angles, with pauses that let the helper fall asleep"""
    for seed, n, rounds in ((1, 4096, 40), (2, 8, 200), (3, 2049, 60), (4, 100000, 6)):
        assert capi.lib().mppb_selftest_sincos_helper(seed, n, rounds) == 0, (seed, n, rounds)
