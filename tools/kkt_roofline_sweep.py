"""KKT kernel alone (one Riccati factorisation + solve per instance, first IPM iteration of a cold solve), CUDA events,
L2 flushed between launches: algorithmic GB/s (SURVEY 8d bytes) against the measured HBM peak.
 (a) batch sweep at N=50 (cfg 2)   (b) BASELINE configs[4]: horizon sweep N in {20,50,100,200} at B=2048 (cfg 5)."""
import sys, json, os; sys.path.insert(0, '.')
import numpy as np
from mpc_local_planner_b200 import capi, configs
peak = 6650.0
try: peak = float(json.load(open('MEASURED_PEAKS.json'))['hbm_gbs'])
except Exception: pass
def run(cid, B, n=None):
    cfg = configs.config_for(cid, n=n, tol=1e-6)
    base = configs.generate(cid, min(B, 2048), n=n)
    rep = (B + 2047) // 2048
    tile = lambda a: np.concatenate([a] * rep)[:B]
    data = dict(x0=tile(base['x0']), xf=tile(base['xf']), u_prev=tile(base['u_prev']), u_prev_dt=base['u_prev_dt'],
                obstacles=tuple(tile(a) for a in base['obstacles']),
                viapoints=tuple(tile(a) for a in base['viapoints']) if base['viapoints'] is not None else None)
    s = capi.BatchSolver(cfg, B)
    s.upload(data['x0'], data['xf'], data['u_prev'], data['u_prev_dt'], data['obstacles'], data['viapoints'])
    s.run_phase(capi.PHASE_INIT); s.run_phase(capi.PHASE_ASSOCIATE); s.run_phase(capi.PHASE_EVAL)
    ms = s.time_phase(capi.PHASE_KKT, reps=10, flush_l2=True)
    N = cfg.n; words = 46 + 4 + 3 * cfg.variable_dt
    byts = B * 8 * (words * (N - 1) + 12)
    gbs = byts / ms / 1e6
    print("cfg %d N %3d B %6d  kkt %.4f ms  algorithmic %.1f MB  %.0f GB/s  = %.1f %% of HBM peak (%.0f GB/s)" % (cid, N, B, ms, byts / 1e6, gbs, 100 * gbs / peak, peak), flush=True)
    s.close()
if "--quick" in sys.argv:
    for B in (1024, 4096): run(2, B)
    run(3, 1024)
    sys.exit(0)
print("# KKT kernel alone, batch sweep, cfg 2 (N=50)")
for B in (1024, 2048, 4096, 8192, 16384, 32768, 65536): run(2, B)
print("# BASELINE configs[4]: horizon sweep at B=2048 (cfg 5)")
for n in (20, 50, 100, 200): run(5, 2048, n)
print("# horizon sweep at B=16384")
for n in (20, 50, 100, 200): run(5, 16384, n)
