import sys; sys.path.insert(0,'.')
import numpy as np
from mpc_local_planner_b200 import configs, capi
from oracle import oracle_py as orc
B=32
cfg=configs.cfg3(tol=1e-8); data=configs.generate(3,B); data['obstacles'][0][:]=0
out=orc.step_batch(cfg,data,n_threads=8)
st=out['status']; it=out['iters']
print("noobs conv",(st==0).sum(),"maxit",(st==1).sum(),"numerr",(st==2).sum(),"iters",it.mean(), "dt",out['dt'][:10].round(3), "T", (out['dt']*79)[:10].round(1), "dist", np.linalg.norm(data['xf'][:10,:2],axis=1).round(1))
