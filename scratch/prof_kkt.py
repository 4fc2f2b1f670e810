import sys; sys.path.insert(0,'.')
from mpc_local_planner_b200 import capi, configs
B=int(sys.argv[1]) if len(sys.argv)>1 else 1024
cfg=configs.config_for(2,tol=1e-6); data=configs.generate(2,B)
s=capi.BatchSolver(cfg,B)
s.upload(data["x0"],data["xf"],data["u_prev"],data["u_prev_dt"],data["obstacles"],data["viapoints"])
s.run_phase(capi.PHASE_INIT); s.run_phase(capi.PHASE_ASSOCIATE)
for it in range(6):
    s.run_phase(capi.PHASE_EVAL); s.run_phase(capi.PHASE_KKT); s.run_phase(capi.PHASE_LINESEARCH)
print(s.stats())
