"""Does a fourth CTA per SM pay with phase alignment?  cfg 2 at N = 48 (its image fits four times into an SM's shared memory)."""
import sys; sys.path.insert(0, '.')
import numpy as np
from mpc_local_planner_b200 import capi, configs
n = int(sys.argv[1]) if len(sys.argv) > 1 else 48
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
cfg = configs.cfg2(n=n)
data = configs.generate(2, B)
for cap in (3, 4):
    for sync in (0, 1):
        s = capi.BatchSolver(cfg, B)
        s.set_option(capi.OPT_SM_PHASE_SYNC, sync)
        s.set_option(capi.OPT_CTAS_PER_SM, cap)
        s.upload(data["x0"], data["xf"], data["u_prev"], data["u_prev_dt"], data["obstacles"], data["viapoints"])
        ts = []
        for r in range(5):
            s.flush_l2(); s.stats_reset()
            ts.append(s.solve_resident(cold=True))
        out = s.fetch(); st = s.stats()
        print(f"N {n} B {B} ctas/sm cap {cap} sync {sync}: min {min(ts)*1e3:8.3f} ms median {np.median(ts)*1e3:8.3f} converged {int((out['status']==0).sum())} iters {out['iters'].mean():.1f} "
              f"phase ms {[round(x,3) for x in st['ms']]} gate {st['gate_ms']:.3f}", flush=True)
        s.close()
