#!/usr/bin/env python
"""Generates the golden fixtures of tests/golden/ (committed; this script is the provenance).

The reference ships no golden vectors and cannot be run here, so the known answers are produced by an
INDEPENDENT algorithm -- scipy.optimize SLSQP (and trust-constr for the single-instance config) with numeric Jacobians --
on the restated OCP functions (objective / dynamics defects / inequality rows evaluated through oracle/liboracle.so),
from the same initial guess the solvers use.  The oracle's interior-point solver and the CUDA solver are then both
tested against these fixtures (tests/test_oracle_golden.py, tests/test_gpu_parity.py).

    python tests/golden/make_golden.py            # rewrites tests/golden/*.json (a few minutes)
"""
import json
import os
import sys
import time

import numpy as np
from scipy.optimize import NonlinearConstraint, minimize

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mpc_local_planner_b200 import capi, configs  # noqa: E402
from oracle import oracle_py as orc  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def scipy_solve(cfg, data, b, method="SLSQP"):
    inst = orc.instance_from_batch(cfg, data, b)
    N = inst.N
    inst.init_cold()
    inst.associate()
    inst.L.orc_project_init(orc.C.byref(inst.p), inst.ws)
    inst.L.orc_init_controls(orc.C.byref(inst.p), inst.ws)
    inst.init_duals()
    X, U, SC = inst.arr("X"), inst.arr("U"), inst.arr("SCAL")
    idx = []
    for k in range(1, N):
        for i in range(3):
            if k == N - 1 and cfg.xf_fixed[i]:
                continue
            idx.append(("x", i, k))
    for k in range(N - 1):
        for i in range(2):
            idx.append(("u", i, k))
    if cfg.variable_dt:
        idx.append(("dt", 0, 0))

    def setz(z):
        for v, (t, i, k) in zip(z, idx):
            if t == "x":
                X[i, k] = v
            elif t == "u":
                U[i, k] = v
            else:
                SC[capi.SC_DT] = v

    def getz():
        return np.array([X[i, k] if t == "x" else (U[i, k] if t == "u" else SC[capi.SC_DT]) for (t, i, k) in idx])

    act = inst.arr("LAM") > 0

    def fun(z):
        setz(z); inst.eval(); return SC[capi.SC_OBJ]

    def ceq(z):
        setz(z); return inst.defects().ravel().copy()   # raw collocation defects (orc_defect per interval)

    def cin(z):
        setz(z); inst.eval(); return -(inst.arr("G")[act]).copy()

    z0 = getz()
    if method == "SLSQP":
        res = minimize(fun, z0, method="SLSQP", constraints=[{"type": "eq", "fun": ceq}, {"type": "ineq", "fun": cin}],
                       options=dict(maxiter=500, ftol=1e-12))
    else:
        res = minimize(fun, z0, method="trust-constr",
                       constraints=[NonlinearConstraint(ceq, 0, 0), NonlinearConstraint(cin, 0, np.inf)],
                       options=dict(maxiter=3000, gtol=1e-9, xtol=1e-12))
    setz(res.x)
    ce = float(np.abs(ceq(res.x)).max())
    ci = float(min(cin(res.x).min(), 0))
    return dict(f=float(res.fun), ceq=ce, cin=ci, nit=int(res.nit), U=U[:, :N - 1].T.copy().tolist(),
                dt=float(SC[capi.SC_DT]), xN=X[:, N - 1].tolist())


def annotate(cfg, data, b, r):
    """Adds the oracle's own optimum of the instance: local methods on a non-convex problem may end in different local optima;
    `agree` marks the fixtures on which SLSQP and the oracle found the same one (the others record that the oracle's is not worse)."""
    inst = orc.instance_from_batch(cfg, data, b)
    u, x, res = inst.step()
    r["oracle_status"] = int(res.status)
    r["f_oracle"] = float(res.objective)
    r["dt_oracle"] = float(res.dt)
    r["agree"] = bool(res.status == 0 and abs(res.objective - r["f"]) <= 1e-6 * max(1.0, abs(r["f"])))
    if not r["agree"]:
        r["U_oracle"] = u[:-1].tolist()
    return r


def main():
    out = {}
    # config 1 / scenario G1 (reference's only fixed scenario): two independent scipy algorithms
    cfg = configs.cfg1(tol=1e-8)
    data = configs.g1_instance()
    t = time.time()
    a = scipy_solve(cfg, data, 0, "SLSQP")
    bb = scipy_solve(cfg, data, 0, "trust-constr")
    print("G1 SLSQP dt %.8f trust-constr dt %.8f  |du| %.2e  (%.0fs)" % (
        a["dt"], bb["dt"], np.abs(np.array(a["U"]) - np.array(bb["U"])).max(), time.time() - t))
    out["g1"] = dict(config_id=1, slsqp=a, trust_constr=bb)
    json.dump(out["g1"], open(os.path.join(HERE, "g1.json"), "w"), indent=1)
    # config 2: a handful of seeded instances (obstacles + rate limits active)
    cfg = configs.cfg2(tol=1e-8)
    data = configs.generate(2, 64)
    ref = orc.step_batch(cfg, data, n_threads=4)
    sel = [b for b in range(64) if ref["status"][b] == 0][:14]
    rows = []
    for b in sel:
        t = time.time()
        r = scipy_solve(cfg, data, b)
        r["instance"] = b
        if r["ceq"] < 1e-8 and r["cin"] > -1e-8 and r["nit"] < 500:
            rows.append(annotate(cfg, data, b, r))
        print("cfg2 inst %d f %.6f ceq %.1e cin %.1e nit %d agree %s (%.0fs)" % (b, r["f"], r["ceq"], r["cin"], r["nit"], r.get("agree"), time.time() - t), flush=True)
        if len(rows) >= 10:
            break
    json.dump(dict(config_id=2, instances=rows), open(os.path.join(HERE, "slsqp_cfg2.json"), "w"), indent=1)


def extra():
    """python tests/golden/make_golden.py extra -- fixtures for the via-point objective (cfg 4) and the car-like minimum-time
    problem with the polygon footprint (cfg 3 at N=30, so that SLSQP with numeric Jacobians finishes in minutes)."""
    for cid, n, fname, want in ((4, None, "slsqp_cfg4.json", 6), (3, 30, "slsqp_cfg3_n30.json", 5)):
        cfg = configs.config_for(cid, n=n, tol=1e-8)
        data = configs.generate(cid, 48, n=n)
        ref = orc.step_batch(cfg, data, n_threads=4)
        sel = [b for b in range(48) if ref["status"][b] == 0][:want + 3]
        rows = []
        for b in sel:
            t = time.time()
            r = scipy_solve(cfg, data, b)
            r["instance"] = b
            print("cfg%d inst %d f %.6f ceq %.1e cin %.1e nit %d dt %.6f (%.0fs)" % (cid, b, r["f"], r["ceq"], r["cin"], r["nit"], r["dt"], time.time() - t), flush=True)
            if r["ceq"] < 1e-8 and r["cin"] > -1e-8 and r["nit"] < 500:
                rows.append(annotate(cfg, data, b, r))
            if len(rows) >= want:
                break
        json.dump(dict(config_id=cid, n=n, instances=rows), open(os.path.join(HERE, fname), "w"), indent=1)


def options():
    """python tests/golden/make_golden.py options -- fixtures for non-default options on cfg 2 (fixed dt): midpoint differences,
    and the integral-form cost integrated by the trapezoidal rule.  (With a free dt the integral-form problems have many
    local optima -- dt l(x, u) is indefinite in (x, u, dt) -- and two local methods rarely meet: no fixture.)"""
    only = sys.argv[2] if len(sys.argv) > 2 else None
    for name, make, fname, want in (("midpoint", lambda: _with(configs.cfg2(tol=1e-8), collocation=capi.COLLOC_MIDPOINT), "slsqp_cfg2_midpoint.json", 5),
                                    ("trapezoidal", lambda: configs.cfg2_trapezoidal(tol=1e-8, variable_dt=False), "slsqp_cfg2_trapezoidal.json", 4)):
        if only and name != only:
            continue
        cfg = make()
        data = configs.generate(2, 32)
        ref = orc.step_batch(cfg, data, n_threads=4)
        sel = [b for b in range(32) if ref["status"][b] == 0][:want + 4]
        rows = []
        for b in sel:
            t = time.time()
            r = scipy_solve(cfg, data, b)
            r["instance"] = b
            print("%s inst %d f %.6f ceq %.1e cin %.1e nit %d dt %.6f (%.0fs)" % (name, b, r["f"], r["ceq"], r["cin"], r["nit"], r["dt"], time.time() - t), flush=True)
            if r["ceq"] < 1e-8 and r["cin"] > -1e-8 and r["nit"] < 500:
                rows.append(annotate(cfg, data, b, r))
            if len(rows) >= want:
                break
        json.dump(dict(config_id=2, option=name, instances=rows), open(os.path.join(HERE, fname), "w"), indent=1)


def _with(cfg, **kw):
    for k, v in kw.items():
        setattr(cfg, k, v)
    return cfg


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "options":
        options()
    elif len(sys.argv) > 1 and sys.argv[1] == "extra":
        extra()
    else:
        main()
