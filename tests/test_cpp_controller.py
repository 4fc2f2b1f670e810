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
