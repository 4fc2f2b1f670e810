"""Per-iteration profile of a cfg-2 solve driven phase by phase through the C ABI (serial KKT attempts): unfinished instances,
device time of eval / KKT / line search, regularisation retries and backtracks.  Usage: python tools/iteration_profile.py"""
import sys; sys.path.insert(0,'.')
import numpy as np
from mpc_local_planner_b200 import capi, configs
B=1024
cfg=configs.config_for(2,tol=1e-6); data=configs.generate(2,B)
s=capi.BatchSolver(cfg,B)
s.upload(data["x0"],data["xf"],data["u_prev"],data["u_prev_dt"],data["obstacles"],data["viapoints"])
s.run_phase(capi.PHASE_INIT); s.run_phase(capi.PHASE_ASSOCIATE)
prev=s.ws_read(capi.F_SCAL)
rows=[]
for it in range(100):
    s.stats_reset()
    s.run_phase(capi.PHASE_EVAL); s.run_phase(capi.PHASE_KKT); s.run_phase(capi.PHASE_LINESEARCH)
    st=s.stats(); sc=s.ws_read(capi.F_SCAL)
    act=prev[:,capi.SC_STATUS]<0
    dn=(sc[:,capi.SC_NREG]-prev[:,capi.SC_NREG])[act]; db=(sc[:,capi.SC_NBT]-prev[:,capi.SC_NBT])[act]
    rows.append((it,int(act.sum()),st['ms'][2]*1e3,st['ms'][3]*1e3,st['ms'][4]*1e3, dn.max() if act.any() else 0, dn.mean() if act.any() else 0, db.max() if act.any() else 0, db.mean() if act.any() else 0, (db>=1).mean() if act.any() else 0))
    prev=sc
    if not (sc[:,capi.SC_STATUS]<0).any(): break
print("it active eval_us kkt_us ls_us max_extra_sweeps mean_extra max_bt mean_bt frac_bt")
for r in rows[::3]: print("%3d %4d %6.1f %6.1f %6.1f %d %.3f %d %.3f %.3f"%r)
print("totals ms: eval %.2f kkt %.2f ls %.2f"%(sum(r[2] for r in rows)/1e3,sum(r[3] for r in rows)/1e3,sum(r[4] for r in rows)/1e3))
