"""GPU parity of the collision-grid construction: the table built on the device by
mppb_build_collision_grid (through the C ABI, mppb_ptg_create_diffdrive_c_on) equals, entry by entry, the
one the checker builds with DiffDriveCollisionGridBased::internal_initialize — the restatement and the
reference's own DiffDriveCollisionGridBased.cpp (the `orc` fixture runs both). Bar: bit-exact."""
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ACK_X = np.float32([-0.10, 0.00, 0.3, 0.3, 0.0, -0.10]).astype(np.float64)
ACK_Y = np.float32([0.15, 0.15, 0.1, -0.1, -0.15, -0.15]).astype(np.float64)


def _same_table(gp, go):
    for key in ("xmin", "ymin", "res", "nx", "ny"):
        assert gp[key] == go[key], key
    assert np.array_equal(gp["offsets"], go["offsets"])
    cnt = np.diff(go["offsets"])
    cell = np.repeat(np.arange(len(cnt)), cnt)
    order = np.lexsort((go["ks"], cell))  # checker lists are in insertion order; the product's are sorted by k
    assert np.array_equal(go["ks"][order], gp["ks"])
    assert np.array_equal(go["ds"][order], gp["ds"])
    assert np.array_equal(go["ks"], gp["ks"])  # ... which IS the insertion order of the k-outermost loop


def test_ackermann_grids_built_on_gpu_equal_checker(capi, orc, oracle_ackermann):
    ctx = capi.Context(0)
    t0 = time.perf_counter()
    ptgs = capi.ackermann_ptgs(ctx=ctx)
    t_gpu = time.perf_counter() - t0
    assert ctx.launch_count >= 8  # 4 kernels per PTG
    for p, o in zip(ptgs, oracle_ackermann):
        _same_table(p.grid(), o.grid())
    print("both Ackermann PTGs incl. GPU grid build: %.1f ms" % (1e3 * t_gpu))
    ctx.close()


def test_gpu_built_equals_host_built(capi, product_ackermann):
    ctx = capi.Context(0)
    for p, q in zip(capi.ackermann_ptgs(ctx=ctx), product_ackermann):
        gp, gq = p.grid(), q.grid()
        for key in ("offsets", "ks", "ds"):
            assert np.array_equal(gp[key], gq[key])
    ctx.close()


@pytest.mark.parametrize("name,args", [
    # 320x320 cells = 410 KB of cell minima: three shared-memory bands per path
    ("multi_band", dict(num_paths=31, ref_distance=8.0, resolution=0.05, v=1.0, w=np.deg2rad(40.0), K=+1.0,
                        sx=ACK_X, sy=ACK_Y)),
    # coarse grid, backward paths, a triangle
    ("triangle_backward", dict(num_paths=15, ref_distance=3.0, resolution=0.10, v=0.7, w=np.deg2rad(90.0), K=-1.0,
                               sx=[0.4, -0.2, -0.2], sy=[0.0, 0.25, -0.25])),
    # a 12-gon larger than the turning radius (cells inside the shape at step 0, grid border clipping)
    ("dodecagon", dict(num_paths=21, ref_distance=2.0, resolution=0.05, v=0.5, w=np.deg2rad(120.0), K=+1.0,
                       sx=(0.6 * np.cos(np.arange(12) * np.pi / 6)).tolist(),
                       sy=(0.45 * np.sin(np.arange(12) * np.pi / 6)).tolist())),
])
def test_other_geometries(capi, orc, name, args):
    ctx = capi.Context(0)
    p = capi.PTG.diffdrive_c(args["num_paths"], args["ref_distance"], args["resolution"], args["v"], args["w"],
                             args["K"], args["sx"], args["sy"], ctx=ctx)
    o = orc.diffdrive_c(args["num_paths"], args["ref_distance"], args["resolution"], args["v"], args["w"],
                        args["K"], args["sx"], args["sy"])
    _same_table(p.grid(), o.grid())
    assert len(p.grid()["ks"]) > 1000
    ctx.close()
