"""The C++ surface of the mpp:: mirror, compiled: tests/cpp/test_mpp_api.cpp (user CostEvaluator subclass on the host
path, progress callback, time budget, class registry, single-call seams) and a plug-in module loaded by
path-planner-cli --plugins (tests/cpp/plugin_planner.cpp). Also the C-ABI seam of transform_pc_square_clipping against
the CPU checker."""
import os
import subprocess

import numpy as np
import pytest

from mrpt_path_planning_b200 import build as _build, workloads as wl

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SHARE = os.path.join(HERE, "golden", "share")
OBS = os.path.join(HERE, "golden", "obstacles_01.txt")
OUT = os.path.join(HERE, "_build")
HOST_INC = os.path.join(ROOT, "mrpt_path_planning_b200", "csrc", "host")


def _compile(src, out, extra):
    os.makedirs(OUT, exist_ok=True)
    _build.build_library()
    deps = [src, os.path.join(HOST_INC, "mpp.h"), os.path.join(ROOT, "include", "mppb.h"), _build.LIB_PATH]
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-ffp-contract=off", "-I" + HOST_INC, src, "-o", out,
                               "-L" + _build.LIB_DIR, "-lmpp_b200", "-Wl,-rpath," + _build.LIB_DIR, "-pthread", "-ldl"] + extra)
    return out


def test_compiled_cpp_api():
    exe = _compile(os.path.join(HERE, "cpp", "test_mpp_api.cpp"), os.path.join(OUT, "test_mpp_api"), [])
    r = subprocess.run([exe, os.path.join(SHARE, "ptgs_ackermann_vehicle.ini"), OBS], stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.strip().startswith("OK user_evaluator_calls=")


def test_cli_loads_a_plugin_planner(tmp_path):
    """--plugins loads a shared object whose initialiser registers a Planner class; --planner instantiates it by name
    (path-planner-cli.cpp:256-265, 452-461). Same plan as with the built-in class."""
    plug = _compile(os.path.join(HERE, "cpp", "plugin_planner.cpp"), os.path.join(OUT, "libtestplugin.so"), ["-shared", "-fPIC"])
    base = [_build.CLI_PATH, "-s", "[0.5 0 0]", "-g", "[4 2.5 45]", "-c", os.path.join(SHARE, "ptgs_ackermann_vehicle.ini"),
            "--obstacles", OBS, "--planner-parameters", os.path.join(SHARE, "mvsim-demo-astar-planner-params.yaml")]
    r0 = subprocess.run(base, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    r1 = subprocess.run(base + ["--plugins", plug, "--planner", "testplugin::LoggingPlanner"], stdout=subprocess.PIPE,
                        stderr=subprocess.PIPE, text=True, timeout=300)
    assert r0.returncode == 0 and r1.returncode == 0, r1.stderr
    assert "[plugin] testplugin::LoggingPlanner::plan" in r1.stdout

    def facts(out):
        return [ln for ln in out.splitlines() if ln.startswith(("Success:", "Plan has", "Path cost:"))]

    assert facts(r0.stdout) == facts(r1.stdout) and len(facts(r0.stdout)) == 3
    # an unknown class name, and a module that does not exist, are reported like the reference does
    r2 = subprocess.run(base + ["--planner", "no::SuchPlanner"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r2.returncode == 1 and "does not seem to be a known C++ class" in r2.stderr
    r3 = subprocess.run(base + ["--plugins", "/nonexistent/libx.so"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r3.returncode == 1 and "Could not load plugins" in r3.stderr


def test_transform_pc_square_clipping_seam_equals_checker(capi, orc):
    """mppb_transform_pc_square_clipping == the reference's function on the same cloud and pose: same points, same
    order, same float bits (transform_pc_square_clipping.cpp:26-42)."""
    ctx = capi.Context(0)
    xs, ys = wl.uniform_cloud(50000, 40.0, seed=5)
    rng = np.random.default_rng(6)
    for _ in range(8):
        pose = [rng.uniform(0, 40), rng.uniform(0, 40), rng.uniform(-np.pi, np.pi)]
        clip = float(rng.choice([0.5, 5.0, 12.0, 100.0]))
        gx, gy = ctx.transform_pc_square_clipping(xs, ys, pose, clip)
        rx, ry = orc.local_obstacles(xs, ys, pose, clip)
        assert np.array_equal(gx.view(np.uint32), np.asarray(rx, np.float32).view(np.uint32))
        assert np.array_equal(gy.view(np.uint32), np.asarray(ry, np.float32).view(np.uint32))
    gx, gy = ctx.transform_pc_square_clipping(xs[:0], ys[:0], [0, 0, 0], 5.0)
    assert gx.size == 0
    ctx.close()
