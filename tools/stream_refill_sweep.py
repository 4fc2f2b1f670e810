"""Refill cadence of the streaming pool (MPCB200_OPT_STREAM_REFILL_EVERY) on cfg 2: device time and pool iterations for a
queue of 16 x 1024 instances.  Usage: python tools/stream_refill_sweep.py [pool] [queue]"""
import sys, time; sys.path.insert(0, '.')
import numpy as np
from mpc_local_planner_b200 import capi, configs
pool = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
total = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
cfg = configs.config_for(2, tol=1e-6)
base = configs.generate(2, 1024)
rep = (total + 1023) // 1024
tile = lambda a: np.ascontiguousarray(np.concatenate([a] * rep)[:total])
q = dict(x0=tile(base["x0"]), xf=tile(base["xf"]), u_prev=tile(base["u_prev"]), obstacles=tuple(tile(a) for a in base["obstacles"]))
s = capi.BatchSolver(cfg, pool)
s.solve_stream(q["x0"], q["xf"], q["u_prev"], base["u_prev_dt"], q["obstacles"])
ref = None
for every in (1, 2, 3, 4, 2, 1):
    s.set_option(capi.OPT_STREAM_REFILL_EVERY, every)
    s.stats_reset()
    t = time.time()
    out = s.solve_stream(q["x0"], q["xf"], q["u_prev"], base["u_prev_dt"], q["obstacles"])
    wall = time.time() - t
    st = s.stats()
    conv = int((out["status"] == 0).sum())
    if ref is None:
        ref = out
    same = np.array_equal(ref["u_seq"], out["u_seq"]) and np.array_equal(ref["status"], out["status"])
    print("refill every %d: device %.1f ms wall %.1f ms -> %.0f converged/s device, %.0f wall; pool iterations %d (expected %.0f); identical results %s" % (
        every, out["solve_time_s"] * 1e3, wall * 1e3, conv / out["solve_time_s"], conv / wall, st["launches"][2], total * out["iters"].mean() / pool, same), flush=True)
s.close()
