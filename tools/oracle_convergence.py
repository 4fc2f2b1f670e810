"""Convergence statistics of the CPU oracle on a BASELINE configuration: python tools/oracle_convergence.py <cfg id> <instances>"""
import sys; sys.path.insert(0,'.')
import numpy as np, time
from mpc_local_planner_b200 import configs, capi
from oracle import oracle_py as orc
cid=int(sys.argv[1]); B=int(sys.argv[2])
cfg = configs.config_for(cid, tol=1e-6)
data = configs.generate(cid, B)
t=time.time(); out = orc.step_batch(cfg, data, n_threads=4); el=time.time()-t
st=out['status']; it=out['iters']
print("cfg",cid,"conv %d/%d"%((st==0).sum(),B),"maxit",(st==1).sum(),"numerr",(st==2).sum(),"total iters",it.sum(),"iters(conv) med %d"%(np.median(it[st==0]) if (st==0).any() else -1),"iters(fail) mean %.1f"%(it[st!=0].mean() if (st!=0).any() else 0),"time %.2fs"%el)
