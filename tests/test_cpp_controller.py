"""The C++ mirror of mpc_local_planner::Controller (include/mpcb200_controller.hpp) driving the reference's only fixed
scenario (test_mpc_optim_node): builds on any box, runs on the GPU box."""
import os
import re
import subprocess

import pytest

from tests.conftest import ROOT, has_gpu

DEMO = os.path.join(ROOT, "examples", "test_mpc_optim_demo")


def _build(cuda_lib):
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "examples"), "-B", "test_mpc_optim_demo"], stdout=subprocess.DEVNULL)


@pytest.mark.skipif(has_gpu(), reason="box has a GPU")
def test_demo_builds_and_fails_loudly_without_gpu(cuda_lib):
    _build(cuda_lib)
    p = subprocess.run([DEMO, "1"], capture_output=True, text=True)
    assert p.returncode == 2
    assert "no CPU fallback" in p.stderr


@pytest.mark.gpu
def test_demo_reproduces_g1_known_answer(cuda_lib):
    """Controller::configure + 3x Controller::step on scenario G1: dt* = 0.71287734, u0* = (0.4, 0.3), x_f reached."""
    _build(cuda_lib)
    p = subprocess.run([DEMO, "3"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    lines = [l for l in p.stdout.splitlines() if l.startswith("step")]
    assert len(lines) == 3
    for l in lines:
        m = re.search(r"ok 1 status 0 iters (\d+) dt ([\d.]+) u0 ([\d.]+) ([\d.]+) xN ([\d.\-e]+) ([\d.\-e]+) ([\d.\-e]+)", l)
        assert m, l
        assert abs(float(m.group(2)) - 0.71287734) < 5e-8
        assert abs(float(m.group(3)) - 0.4) < 1e-6 and abs(float(m.group(4)) - 0.3) < 1e-6
        assert abs(float(m.group(5)) - 5.0) < 1e-8 and abs(float(m.group(6)) - 2.0) < 1e-8


@pytest.mark.gpu
def test_demo_grid_adaptation_follows_the_optimal_dt(cuda_lib):
    """grid/variable_grid/grid_adaptation as test_mpc_optim_node.yaml:54-58 enables it: the first solve keeps grid_size_ref
    = 20 (dt* = 0.713 > 0.33), then the horizon grows by one grid point per step until the optimal dt is inside the
    hysteresis band dt_ref (1 +- 0.1); the horizon time stays near the minimum time of the problem throughout."""
    _build(cuda_lib)
    p = subprocess.run([DEMO, "30", "1"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    rows = [re.search(r"status (\d+) iters (\d+) dt ([\d.]+) .* n (\d+)$", l) for l in p.stdout.splitlines() if l.startswith("step")]
    assert len(rows) == 30 and all(rows)
    n = [int(m.group(4)) for m in rows]; dt = [float(m.group(3)) for m in rows]; st = [int(m.group(1)) for m in rows]
    assert all(s == 0 for s in st)
    assert n[0] == 20 and abs(dt[0] - 0.71287734) < 5e-8
    for i in range(1, 30):
        want = n[i - 1] + 1 if dt[i - 1] > 0.33 else (n[i - 1] - 1 if dt[i - 1] < 0.27 else n[i - 1])
        assert n[i] == want
    assert 0.27 <= dt[-1] <= 0.33 and n[-1] == n[-2] > 40
    T = [(k - 1) * d for k, d in zip(n, dt)]
    assert max(T) - min(T) < 0.03 * T[0] and T[-1] <= T[0] + 1e-9
