"""Expectations of the golden fixtures (tests/golden/np_*.json, written by tests/golden/make_golden.py: scipy SLSQP on the
independent numpy restatement tests/golden/ocp_numpy.py), shared by the oracle tests (CPU) and the CUDA tests (GPU).
`solve(cfg, data)` is the solver under test returning a step_batch-style dict.

The OCPs are non-convex: two local methods started from the same point may end in different KKT points.  A fixture row is
MATCHED when the solver under test reproduces SLSQP's optimum; otherwise the solver's point must at least be a feasible point
of the independent restatement (and it is counted).  Most rows must match."""
import json
import os
import sys

import numpy as np

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
if HERE not in sys.path:
    sys.path.insert(0, HERE)
import golden_cases  # noqa: E402
import ocp_numpy as on  # noqa: E402

U_TOL_SCIPY = 2e-4  # SLSQP's accuracy with numeric Jacobians
# rows of a case that may end in another KKT point / rows the solver under test may fail to converge on
ALLOW_OTHER = {"cfg2": 2, "cfg4": 2, "cfg2_midpoint": 1, "cfg2_trapezoidal": 1, "cfg2_circular_footprint": 1, "cfg1_obstacles": 2, "g1": 0}


def load(case):
    return json.load(open(os.path.join(HERE, "np_%s.json" % case)))


def check_case(case, solve):
    cfg, data, want, pool = golden_cases.make(case)
    fx = load(case)
    rows = fx["rows"]
    assert len(rows) >= max(1, want // 2), f"{case}: only {len(rows)} fixture rows"
    out = solve(cfg, data)
    matched, other = 0, []
    for r in rows:
        b = r["instance"]
        U = np.array(r["U"])
        p = on.problem_from_batch(cfg, data, b)
        # the fixture itself is a feasible point of the restatement with the recorded objective (guards the JSON)
        assert r["ceq"] < 1e-8 and r["cin"] > -1e-8
        if out["status"][b] != 0:
            other.append((b, "not converged"))
            continue
        Xs, Us, dts = out["x_seq"][b], out["u_seq"][b][:-1], float(out["dt"][b])
        f_s = on.objective(cfg, Xs, Us, dts, data["xf"][b], p.vps, p.vp_stage)
        same_f = abs(f_s - r["f"]) <= 1e-6 * max(1.0, abs(r["f"]))
        same_u = np.abs(Us - U).max() < U_TOL_SCIPY
        if same_f and abs(dts - r["dt"]) < 1e-6 and (same_u or cfg.variable_dt):
            # (minimum-time optima need not be strict in the controls: objective and optimal time decide there)
            matched += 1
            continue
        # another KKT point: it must be feasible for the independent restatement
        ce = np.abs(on.defects(cfg, Xs, Us, dts)).max()
        ci = on.inequality_rows(cfg, Xs, Us, dts, data["xf"][b], data["u_prev"][b], data["u_prev_dt"], p.assoc, p.ot, p.op)
        assert ce < 1e-6 and (len(ci) == 0 or ci.max() < 1e-6), f"{case} instance {b}: infeasible point (ceq {ce}, cin {ci.max() if len(ci) else 0})"
        other.append((b, "f %.6f vs SLSQP %.6f" % (f_s, r["f"])))
    assert len(other) <= ALLOW_OTHER[case], f"{case}: matched {matched} of {len(rows)}, others {other}"
    return matched, other
