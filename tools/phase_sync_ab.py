"""A/B of the execution options of the persistent solve kernel (results are identical in every mode; checked):
MPCB200_OPT_SM_PHASE_SYNC (phase alignment of the CTAs that share an SM: 0 off, 1 three gates, 2 two gates) x
MPCB200_OPT_ORDER_BY_HISTORY (queue longest-first by the previous solve's iteration counts: 0 / 1).
Device time of cold batch solves, L2 flushed before each (the first solve of a handle has no history).
usage: python tools/phase_sync_ab.py [cfg=2] [B=1024] [reps=7] [first_instance=0]"""
import sys; sys.path.insert(0, '.')
import numpy as np
from mpc_local_planner_b200 import capi, configs
cid = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 7
first = int(sys.argv[4]) if len(sys.argv) > 4 else 0
cfg = configs.config_for(cid)
data = configs.generate(cid, B, first=first)
res = {}
for sync, order in ((0, 0), (1, 0), (2, 0), (1, 1)):
    s = capi.BatchSolver(cfg, B)
    s.set_option(capi.OPT_SM_PHASE_SYNC, sync)
    s.set_option(capi.OPT_ORDER_BY_HISTORY, order)
    s.upload(data["x0"], data["xf"], data["u_prev"], data["u_prev_dt"], data["obstacles"], data["viapoints"])
    ts = []
    for r in range(reps + 1):
        s.flush_l2(); s.stats_reset()
        t = s.solve_resident(cold=True)
        if r > 0: ts.append(t)
    out = s.fetch(); st = s.stats()
    res[(sync, order)] = out
    print(f"cfg {cid} B {B} phase sync {sync} order by history {order}: min {min(ts)*1e3:8.3f} ms median {np.median(ts)*1e3:8.3f}  converged {int((out['status']==0).sum())} "
          f"phase ms {[round(x,3) for x in st['ms']]} gate {st['gate_ms']:.3f}", flush=True)
    s.close()
a = res[(0, 0)]
print("identical results:", all(bool((a["u_seq"] == o["u_seq"]).all() and (a["iters"] == o["iters"]).all()) for o in res.values()))
