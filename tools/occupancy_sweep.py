"""Per-call phase times of the solve kernel against the number of CTAs resident per SM (batch = 148 x CTAs: one wave).
usage: python tools/occupancy_sweep.py [cfg=2]"""
import sys; sys.path.insert(0, '.')
import numpy as np
from mpc_local_planner_b200 import capi, configs
cid = int(sys.argv[1]) if len(sys.argv) > 1 else 2
cfg = configs.config_for(cid)
for per_sm in (1, 2, 3, 4):
    B = 148 * per_sm
    data = configs.generate(cid, B)
    s = capi.BatchSolver(cfg, B)
    s.set_option(capi.OPT_CTAS_PER_SM, per_sm)
    s.upload(data["x0"], data["xf"], data["u_prev"], data["u_prev_dt"], data["obstacles"], data["viapoints"])
    ts = []
    for r in range(3):
        s.flush_l2(); s.stats_reset(); ts.append(s.solve_resident(cold=True))
    out = s.fetch(); st = s.stats()
    calls = st["kkt_instances"] / B   # iterations per CTA (one instance each)
    ms = st["ms"]
    print(f"cfg {cid} CTAs/SM {per_sm} B {B}: {min(ts)*1e3:7.3f} ms  iters max {out['iters'].max()} mean {out['iters'].mean():.1f}  "
          f"per call: eval {ms[2]/calls*1e3:6.1f} us  kkt {ms[3]/calls*1e3:6.1f} us  linesearch {ms[4]/calls*1e3:6.1f} us  "
          f"(init {ms[0]*1e3:.0f} us, associate {ms[1]*1e3:.0f} us per instance)", flush=True)
    s.close()
