import sys; sys.path.insert(0,'.')
import numpy as np
from mpc_local_planner_b200 import configs, capi
from oracle import oracle_py as orc
b=int(sys.argv[1])
cfg=configs.cfg2(tol=1e-8); data=configs.generate(2,b+1)
inst=orc.instance_from_batch(cfg,data,b)
u,x,r=inst.step()
print(r.status, r.iters)
for n in ('X','U','NU','S','LAM','KKT','STEP'):
    a=inst.arr(n); print(n, "nan",np.isnan(a).sum(),"inf",np.isinf(a).sum(),"absmax",np.nanmax(np.abs(a)), "min", np.nanmin(a))
print(inst.arr('SCAL'))
