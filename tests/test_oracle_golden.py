"""The oracle's interior-point solver against the committed golden fixtures (tests/golden/np_*.json, produced by
tests/golden/make_golden.py: scipy SLSQP on tests/golden/ocp_numpy.py, a numpy restatement of the OCP written from the
reference's source lines that neither loads nor calls the oracle)."""
import json
import os

import numpy as np
import pytest

from mpc_local_planner_b200 import configs

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
# scipy with numeric Jacobians is itself only ~1e-5 accurate in u; the objective agrees much tighter
U_TOL_SCIPY = 2e-4


def test_g1_known_answer(orc):
    """Scenario G1 = the reference's only fixed scenario (src/test_mpc_optim_node.cpp:67-69,105-106): minimum-time
    unicycle, N=20.  Known answer (SLSQP on the numpy restatement): dt* = 0.71287734, T* = 13.54467 s, u0* = (0.4, 0.3);
    the three obstacles are inactive (inactive rows must be exact no-ops)."""
    import golden_checks as gc
    g = gc.load("g1")["rows"][0]
    assert g["dt"] == pytest.approx(0.71287734, abs=2e-7) and g["f"] == pytest.approx(13.54467, abs=1e-5)
    for _ in (0,):
        cfg = configs.cfg1(tol=1e-9)
        data = configs.g1_instance()
        out = orc.step_batch(cfg, data)
        assert out["status"][0] == 0
        assert out["dt"][0] == pytest.approx(0.71287734, abs=2e-8)
        assert out["dt"][0] == pytest.approx(g["dt"], abs=1e-7)
        assert 19 * out["dt"][0] == pytest.approx(13.54467, abs=1e-5)
        np.testing.assert_allclose(out["u_seq"][0][0], [0.4, 0.3], atol=1e-6)
        np.testing.assert_allclose(out["u_seq"][0][:-1], np.array(g["U"]), atol=U_TOL_SCIPY)
        np.testing.assert_allclose(out["x_seq"][0][-1], [5.0, 2.0, 0.0], atol=1e-9)
    # same optimum without the (inactive) obstacles
    d2 = dict(data); d2["obstacles"] = None
    out2 = orc.step_batch(cfg, d2)
    assert out2["dt"][0] == pytest.approx(out["dt"][0], abs=1e-9)


@pytest.mark.parametrize("case", ["cfg2", "cfg4", "cfg2_midpoint", "cfg2_trapezoidal", "cfg2_circular_footprint", "cfg1_obstacles"])
def test_golden_cases(orc, case):
    """The oracle's interior-point solver against SLSQP on the independent numpy restatement (tests/golden/ocp_numpy.py), from
    the reference's cold initial guess."""
    import golden_checks as gc
    matched, other = gc.check_case(case, lambda cfg, data: orc.step_batch(cfg, data, n_threads=4))
    assert matched >= 1


def test_converged_solutions_are_feasible_kkt_points(orc):
    """Independent of any solver: at the returned point the reference-form defects vanish, all rows hold, the bounds hold."""
    import ctypes as C
    cfg = configs.cfg2(tol=1e-9)
    data = configs.generate(2, 12)
    L = orc.lib()
    for b in range(12):
        inst = orc.instance_from_batch(cfg, data, b)
        u, x, res = inst.step()
        if res.status != 0:
            continue
        X = inst.arr("X"); U = inst.arr("U")
        for k in range(cfg.n - 1):
            e = np.zeros(3)
            x1 = np.ascontiguousarray(X[:, k]); x2 = np.ascontiguousarray(X[:, k + 1]); uk = np.ascontiguousarray(U[:, k])
            L.orc_defect_reference(C.byref(inst.cfg), x1.ctypes.data_as(C.POINTER(C.c_double)), uk.ctypes.data_as(C.POINTER(C.c_double)),
                                   x2.ctypes.data_as(C.POINTER(C.c_double)), cfg.dt_ref, e.ctypes.data_as(C.POINTER(C.c_double)))
            assert np.abs(e).max() < 1e-7
        assert (U[0, :-1] <= 0.4 + 1e-9).all() and (U[0, :-1] >= -0.2 - 1e-9).all()
        assert (np.abs(U[1, :-1]) <= 0.3 + 1e-9).all()
        du = np.diff(np.concatenate([[0.0], U[0, :-1], [0.0]]))
        assert (np.abs(du[1:]) <= 0.2 * cfg.dt_ref + 1e-8).all() and abs(du[0]) <= 0.2 * 0.2 + 1e-8
        inst.eval()
        assert (inst.arr("G") <= 1e-7).all()
