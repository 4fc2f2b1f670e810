import sys; sys.path.insert(0,'.')
import numpy as np, time
from mpc_local_planner_b200 import configs, capi
from oracle import oracle_py as orc
b=int(sys.argv[1])
cfg = configs.cfg2(tol=1e-8); data = configs.generate(2,b+1); data['obstacles'][0][:]=0
inst = orc.instance_from_batch(cfg, data, b)
u,x,r = inst.step()
print(b, r.status, r.iters, "%.2e"%r.kkt_err, "obj %.4f"%r.objective, "reg",r.n_regularised,"bt",r.n_backtracks, u[0], x[-1], r.dt)
