"""Expectations of the scipy fixtures (tests/golden/*.json, written by tests/golden/make_golden.py), shared by the oracle
tests (CPU) and the CUDA tests (GPU).  `out` is a step_batch-style dict.

The OCPs are non-convex: two local methods started from the same guess may end in different local optima.  Every fixture
therefore records whether SLSQP and the oracle agreed when it was generated (`agree`), and for the others the oracle's own
optimum, which is never worse than SLSQP's.  A solver under test must reproduce SLSQP's controls on the agreeing fixtures and
the oracle's on the others."""
import json
import os

import numpy as np

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
U_TOL_SCIPY = 2e-4  # SLSQP's accuracy with numeric Jacobians


def load(name):
    return json.load(open(os.path.join(HERE, name)))


def check_fixed_dt(out, rows, min_rows=5):
    assert len(rows) >= min_rows
    assert sum(r["agree"] for r in rows) >= (len(rows) + 1) // 2
    for r in rows:
        b = r["instance"]
        assert out["status"][b] == 0, f"instance {b} did not converge"
        assert r["f_oracle"] <= r["f"] + 1e-6 * max(1.0, abs(r["f"]))  # recorded at generation: the oracle's optimum is not worse
        U = np.array(r["U"] if r["agree"] else r["U_oracle"])
        assert np.abs(out["u_seq"][b][: U.shape[0]] - U).max() < (U_TOL_SCIPY if r["agree"] else 1e-5)


def check_cfg3_n30(out, rows):
    """Minimum-time problems around obstacles also have non-strict optimal controls: only the optimal time is compared --
    it must equal the oracle's recorded one, which is SLSQP's on most fixtures and never longer."""
    assert len(rows) >= 4
    same = 0
    for r in rows:
        b = r["instance"]
        assert out["status"][b] == 0, f"instance {b} did not converge"
        assert r["dt_oracle"] <= r["dt"] * (1.0 + 1e-6)
        assert abs(out["dt"][b] - r["dt_oracle"]) < 1e-6
        same += abs(r["dt_oracle"] - r["dt"]) < 1e-5
    assert same >= (len(rows) + 1) // 2


def option_config(name, tol=1e-9):
    """the configurations of the option fixtures (tests/golden/make_golden.py options)"""
    from mpc_local_planner_b200 import capi, configs
    if name == "midpoint":
        cfg = configs.cfg2(tol=tol)
        cfg.collocation = capi.COLLOC_MIDPOINT
        return cfg
    assert name == "trapezoidal"
    return configs.cfg2_trapezoidal(tol=tol, variable_dt=False)
