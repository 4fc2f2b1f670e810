"""Streaming (continuous batching) vs batch-at-a-time on cfg 2: device time, iterations run by the pool, per-phase times.
Usage: python tools/stream_profile.py [pool] [queue]"""
import sys, time; sys.path.insert(0, '.')
import numpy as np
from mpc_local_planner_b200 import capi, configs
pool = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
total = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
cfg = configs.config_for(2, tol=1e-6)
base = configs.generate(2, 1024)
rep = (total + 1023) // 1024
tile = lambda a: np.ascontiguousarray(np.concatenate([a] * rep)[:total])
q = dict(x0=tile(base["x0"]), xf=tile(base["xf"]), u_prev=tile(base["u_prev"]), obstacles=tuple(tile(a) for a in base["obstacles"]))
s = capi.BatchSolver(cfg, pool)
for mask in (1 << 3, 0x1f):
    s.set_timing(mask)
    s.solve_stream(q["x0"][:2 * pool], q["xf"][:2 * pool], q["u_prev"][:2 * pool], base["u_prev_dt"], tuple(a[:2 * pool] for a in q["obstacles"]))
    s.stats_reset()
    t = time.time()
    out = s.solve_stream(q["x0"], q["xf"], q["u_prev"], base["u_prev_dt"], q["obstacles"])
    wall = time.time() - t
    st = s.stats()
    conv = int((out["status"] == 0).sum())
    print("timing mask %#x: queue %d pool %d: device %.1f ms wall %.1f ms -> %.0f converged/s (device), mean iters %.1f, pool iterations %d (%.1f us each), launches %d" % (
        mask, total, pool, out["solve_time_s"] * 1e3, wall * 1e3, conv / out["solve_time_s"], out["iters"].mean(), st["launches"][2],
        out["solve_time_s"] * 1e6 / max(st["launches"][2], 1), st["launches_total"]))
    print("    phase ms: init %.2f associate %.2f eval %.2f kkt %.2f linesearch %.2f ; expected pool iterations = queue*mean/pool = %.0f" % (
        *st["ms"], total * out["iters"].mean() / pool))
s.close()
