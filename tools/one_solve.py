"""Two cold solves of cfg 2 (B from argv, default 1024) -- the target of the ncu captures under profiles/."""
import sys; sys.path.insert(0,'.')
from mpc_local_planner_b200 import capi, configs
B=int(sys.argv[1]) if len(sys.argv)>1 else 1024
cfg=configs.config_for(2,tol=1e-6); data=configs.generate(2,B)
s=capi.BatchSolver(cfg,B)
s.upload(data["x0"],data["xf"],data["u_prev"],data["u_prev_dt"],data["obstacles"],data["viapoints"])
for rep in range(2):
    t=s.solve_resident(cold=True)
out=s.fetch(); print("dev ms",t*1e3,"conv",(out['status']==0).sum(), s.stats())
