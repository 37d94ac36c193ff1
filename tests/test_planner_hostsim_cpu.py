"""CPU test of the product planner's HOST logic (mpp::TPS_Astar in libmpp_b200: candidate enumeration, selection,
tentative edges, relaxation, open set, lattice, motion tree, split collision calls, deferred tree insertions, lock-step
batches on the worker pool) with the device entry points answered by the oracle.

libmpp_b200.so is used unmodified; a stand-in for the DEVICE functions of include/mppb.h (tests/hostsim/mppb_sim.cpp,
test infrastructure) is pre-loaded into a child process, where the host code's calls reach it through the PLT. The child
(tests/hostsim/run_plans.py) asserts that every plan equals the oracle planner's motion tree. The kernels' own parity is
the `-m gpu` tests' job; this test says nothing about them."""
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SIM_SRC = os.path.join(HERE, "hostsim", "mppb_sim.cpp")
SIM_LIB = os.path.join(HERE, "_build", "libmppb_sim.so")


def build_sim():
    from oracle import orc

    orc.build()
    odir = os.path.join(ROOT, "oracle", "_build")
    os.makedirs(os.path.dirname(SIM_LIB), exist_ok=True)
    if not os.path.exists(SIM_LIB) or os.path.getmtime(SIM_LIB) < max(
            os.path.getmtime(SIM_SRC), os.path.getmtime(os.path.join(ROOT, "include", "mppb.h"))):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-Wall", SIM_SRC, "-o", SIM_LIB,
                               "-L" + odir, "-loracle", "-Wl,-rpath," + odir])
    return SIM_LIB


import pytest


@pytest.mark.parametrize("expand", [True, False], ids=["device_expansion", "host_phases"])
def test_product_planner_host_logic_equals_oracle_planner(capi, expand):
    """expand=True: collision-grid PTG sets go through mppb_expand_nodes (answered by the stand-in in the reference's
    sequential form) and the host does relaxation only; expand=False (MPPB_NO_EXPAND): the full host phases A, C, D, F."""
    capi.lib()  # the product library is built and loads
    env = dict(os.environ, LD_PRELOAD=build_sim())
    env.pop("MPPB_NO_EXPAND", None)
    if not expand:
        env["MPPB_NO_EXPAND"] = "1"
    r = subprocess.run([sys.executable, os.path.join(HERE, "hostsim", "run_plans.py")], env=env, stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert out["expand"] is expand
    assert out["c2"]["expanded"] > 5 and out["cap_and_point_goal"] and out["batch"]["expansions"] > 50
    assert out["obstacles_changed_between_plans"] is True
    assert out["c1"]["expanded"] > 20 and out["holonomic_batch"] and len(out["parameter_sets"]) == 6
    assert out["real_context_available"] is False or os.path.exists("/dev/nvidiactl")  # no CPU path in the product
