"""Cold batch solves against the cap on resident CTAs per SM: python tools/cap_sweep.py cfg B"""
import sys; sys.path.insert(0, '.')
import numpy as np
from mpc_local_planner_b200 import capi, configs
cid = int(sys.argv[1]); B = int(sys.argv[2])
cfg = configs.config_for(cid)
data = configs.generate(cid, B)
for cap in (1, 2, 3, 4):
    s = capi.BatchSolver(cfg, B)
    s.set_option(capi.OPT_CTAS_PER_SM, cap)
    s.upload(data["x0"], data["xf"], data["u_prev"], data["u_prev_dt"], data["obstacles"], data["viapoints"])
    ts = []
    for r in range(3):
        s.flush_l2(); ts.append(s.solve_resident(cold=True))
    out = s.fetch()
    conv = int((out["status"] == 0).sum())
    print(f"cfg {cid} B {B} cap {cap}: {min(ts)*1e3:8.3f} ms  converged {conv} => {conv/min(ts):,.0f} solves/s", flush=True)
    s.close()
