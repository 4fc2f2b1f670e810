"""Whole-batch parity report GPU vs CPU oracle on the BASELINE configurations (tol 1e-8 so that the optimum is resolved):
status agreement, and over the instances both converge on: |u* - u*_oracle|_inf, |dt* - dt*_oracle|.
Usage: python tools/parity_report.py [B]"""
import sys; sys.path.insert(0, '.')
import numpy as np
from mpc_local_planner_b200 import capi, configs
from oracle import oracle_py as orc

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
for cid in (2, 3, 4, 5):
    cfg = configs.config_for(cid, tol=1e-8)
    data = configs.generate(cid, B)
    s = capi.BatchSolver(cfg, B)
    out = s.step(data["x0"], data["xf"], data["u_prev"], data["u_prev_dt"], data["obstacles"], data["viapoints"])
    s.close()
    ref = orc.step_batch(cfg, data, n_threads=64)
    both = (out["status"] == 0) & (ref["status"] == 0)
    du = np.abs(out["u_seq"][both] - ref["u_seq"][both]).reshape(both.sum(), -1).max(axis=1)
    ddt = np.abs(out["dt"][both] - ref["dt"][both])
    print("cfg %d B %d: converged gpu %d oracle %d both %d | status agreement %.4f | du max %.2e, share < 1e-4: %.4f, median %.1e | ddt max %.1e | iteration count equal on %.3f" % (
        cid, B, (out["status"] == 0).sum(), (ref["status"] == 0).sum(), both.sum(), (out["status"] == ref["status"]).mean(),
        du.max(), (du < 1e-4).mean(), np.median(du), ddt.max(), (out["iters"] == ref["iters"]).mean()), flush=True)
