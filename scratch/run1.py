import sys; sys.path.insert(0,'.')
import numpy as np, time
from mpc_local_planner_b200 import configs, capi
from oracle import oracle_py as orc
cid=int(sys.argv[1]); b=int(sys.argv[2])
cfg = configs.config_for(cid, tol=1e-8)
data = configs.generate(cid, b+1) if cid!=1 else configs.g1_instance()
inst = orc.instance_from_batch(cfg, data, b)
u,x,r = inst.step()
print(b, r.status, r.iters, "%.2e"%r.kkt_err, "obj %.4f"%r.objective, "reg",r.n_regularised,"bt",r.n_backtracks, u[0], x[-1], r.dt)
