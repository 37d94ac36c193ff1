"""CPU tests: the product's host-side PTG tables equal the oracle's, and the C-ABI library
loads and exports every symbol declared in include/mppb.h (no compute calls, no GPU)."""
import ctypes
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(capi):
    hdr = open(os.path.join(ROOT, "include", "mppb.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(mppb_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) > 20
    L = ctypes.CDLL(capi.lib_path())
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, missing
    assert declared <= set(capi._SIGS), sorted(declared - set(capi._SIGS))
    assert b"sm_100a" in capi.lib().mppb_version()


def test_no_cpu_fallback_without_gpu(capi):
    import torch

    if torch.cuda.is_available():
        return
    try:
        capi.Context(0)
    except capi.MppbError as e:
        assert "no usable CUDA device" in str(e)
    else:
        raise AssertionError("context creation must fail without a GPU")


def test_diffdrive_tables_equal_oracle(product_ackermann, oracle_ackermann):
    for p, o in zip(product_ackermann, oracle_ackermann):
        assert p.num_paths == o.num_paths == 121
        assert p.max_radius == o.max_radius
        assert p.step_duration == o.step_duration
        for k in range(0, 121, 7):
            n = p.step_count(k)
            assert n == o.step_count(k)
            tr = o.path(k)
            for s in (0, 1, 2, n // 3, n // 2, n - 2, n - 1):
                pp = p.path_pose(k, s)
                assert np.array_equal(pp, o.path_pose(k, s))
                assert pp[0] == tr[s, 0] and pp[3] == tr[s, 4]
            assert p.init_dist(k) == o.init_dist(k)
        gp, go = p.grid(), o.grid()
        for key in ("xmin", "ymin", "res", "nx", "ny"):
            assert gp[key] == go[key]
        assert np.array_equal(gp["offsets"], go["offsets"])
        cnt = np.diff(go["offsets"])
        cell = np.repeat(np.arange(len(cnt)), cnt)
        order = np.lexsort((go["ks"], cell))  # oracle lists are in insertion order: sort by (cell, k)
        assert np.array_equal(go["ks"][order], gp["ks"])
        assert np.array_equal(go["ds"][order], gp["ds"])


def test_diffdrive_host_queries_equal_oracle(product_ackermann, oracle_ackermann):
    rng = np.random.default_rng(7)
    for p, o in zip(product_ackermann, oracle_ackermann):
        for _ in range(300):
            x, y = rng.uniform(-4, 4, 2)
            assert p.inverse_map(x, y) == o.inverse_map(x, y)
        assert p.inverse_map(1.5, 0.0) == o.inverse_map(1.5, 0.0)
        assert p.inverse_map(-1.5, 0.0) == o.inverse_map(-1.5, 0.0)
        for k in (0, 33, 60, 120):
            for d in (0.0, 0.4, 2.5, 4.999, 5.0, 7.0):
                assert p.step_for_dist(k, d) == o.step_for_dist(k, d)
