import sys; sys.path.insert(0,'.')
import numpy as np, time
from mpc_local_planner_b200 import configs, capi
from oracle import oracle_py as orc
cid=int(sys.argv[1]); B=int(sys.argv[2]); mi=int(sys.argv[3]); tol=float(sys.argv[4]) if len(sys.argv)>4 else 1e-8
cfg = configs.config_for(cid, tol=tol); cfg.max_iter=mi
data = configs.generate(cid, B)
t=time.time(); out = orc.step_batch(cfg, data, n_threads=4); el=time.time()-t
st=out['status']; it=out['iters']
print("cfg",cid,"maxit",mi,"tol",tol,"conv %d/%d"%((st==0).sum(),B),"maxit",(st==1).sum(),"numerr",(st==2).sum(),"iters(conv) mean %.1f med %d p90 %d"%(it[st==0].mean(),np.median(it[st==0]),np.percentile(it[st==0],90)),"time %.2fs"%el)
