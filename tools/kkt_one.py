"""One KKT launch on a prepared batch (for ncu): python tools/kkt_one.py cfg B"""
import sys; sys.path.insert(0, '.')
from mpc_local_planner_b200 import capi, configs
cid = int(sys.argv[1]); B = int(sys.argv[2])
cfg = configs.config_for(cid); data = configs.generate(cid, B)
s = capi.BatchSolver(cfg, B)
s.upload(data["x0"], data["xf"], data["u_prev"], data["u_prev_dt"], data["obstacles"], data["viapoints"])
s.run_phase(capi.PHASE_INIT); s.run_phase(capi.PHASE_ASSOCIATE); s.run_phase(capi.PHASE_EVAL)
print("kkt ms", s.time_phase(capi.PHASE_KKT, reps=3, flush_l2=True))
sc = s.ws_read(capi.F_SCAL); print("defer", sc[:, capi.SC_DEFER].sum(), "delta", sc[:4, capi.SC_DELTA], "nreg", sc[:, capi.SC_NREG].mean())
