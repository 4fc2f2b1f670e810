import sys; sys.path.insert(0,'.')
import numpy as np
from mpc_local_planner_b200 import configs, capi
from oracle import oracle_py as orc
cfg=configs.config_for(2,tol=1e-6); B=192; data=configs.generate(2,B)
tot=[];conv=0;its=0
hist=np.zeros(16,int)
for b in range(B):
    o=orc.instance_from_batch(cfg,data,b); u,x,r=o.step(); conv+=r.status==0; its+=r.iters; tot.append(r.n_regularised)
print("conv",conv,"iters",its,"total extra sweeps",sum(tot),"per iter %.3f"%(sum(tot)/its))
