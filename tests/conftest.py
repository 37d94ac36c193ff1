import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", params=["port", "reference"])
def orc(request):
    """The CPU checker (test infrastructure; never used by the product), in its two forms:
    "port"      = oracle/orc_*.hpp, the line-by-line restatement (liboracle.so);
    "reference" = the reference's own translation units compiled from /root/reference against the
                  MRPT stand-in (oracle/_ref/libmpp_ref.so, built here, shipped prebuilt to the GPU box).
    Every test that takes this fixture runs against both, through the same orc_* interface."""
    if request.param == "port":
        from oracle import orc as _orc

        _orc.lib()
        return _orc
    from oracle import ref as _ref

    if not _ref.available():
        pytest.skip("oracle/_ref/libmpp_ref.so not built and /root/reference absent")
    return _ref.module()


@pytest.fixture(scope="session")
def capi():
    from mrpt_path_planning_b200 import capi as _capi

    _capi.lib()
    return _capi


@pytest.fixture(scope="session")
def oracle_ackermann(orc):
    return orc.ackermann_ptgs()


@pytest.fixture(scope="session")
def product_ackermann(capi):
    return capi.ackermann_ptgs()


@pytest.fixture(scope="session")
def gpu_ctx(capi, product_ackermann):
    """A context on cuda:0 with the Ackermann PTG tables uploaded. Fails loudly without a GPU."""
    ctx = capi.Context(0)
    for p in product_ackermann:
        ctx.add_ptg(p)
    yield ctx
    ctx.close()


def free_dist(out_f32, ptgs):
    """min(init_k, out) in double, as the reference's tp_obstacles_single_path returns it."""
    init = np.concatenate([[p.init_dist(k) for k in range(p.num_paths)] for p in ptgs])
    return np.minimum(init[None, :], out_f32.astype(np.float64))
