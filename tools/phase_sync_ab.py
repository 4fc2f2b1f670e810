"""A/B of MPCB200_OPT_SM_PHASE_SYNC: device time of cold batch solves with and without phase alignment of co-resident CTAs.
usage: python tools/phase_sync_ab.py [cfg=2] [B=1024] [reps=7]"""
import sys; sys.path.insert(0, '.')
import numpy as np
from mpc_local_planner_b200 import capi, configs
cid = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 7
cfg = configs.config_for(cid)
first = int(sys.argv[4]) if len(sys.argv) > 4 else 0
data = configs.generate(cid, B, first=first)
res = {}
for cap in (0, 1):   # here: cap = order-by-history off / on
    for sync in (1,):
        s = capi.BatchSolver(cfg, B)
        s.set_option(capi.OPT_SM_PHASE_SYNC, sync)
        s.set_option(capi.OPT_ORDER_BY_HISTORY, cap)
        s.upload(data["x0"], data["xf"], data["u_prev"], data["u_prev_dt"], data["obstacles"], data["viapoints"])
        ts = []
        for r in range(reps):
            s.flush_l2(); s.stats_reset()
            ts.append(s.solve_resident(cold=True))
        out = s.fetch(); st = s.stats()
        res[(cap, sync)] = out
        print(f"cfg {cid} B {B} ctas/sm cap {cap} sync {sync}: min {min(ts)*1e3:8.3f} ms median {np.median(ts)*1e3:8.3f}  converged {int((out['status']==0).sum())} "
              f"phase ms {[round(x,3) for x in st['ms']]} gate {st['gate_ms']:.3f}", flush=True)
        s.close()
a = res[(0, 1)]
print("identical results:", all(bool((a["u_seq"] == o["u_seq"]).all() and (a["iters"] == o["iters"]).all()) for o in res.values()))
