import sys; sys.path.insert(0,'.')
import numpy as np, time
from mpc_local_planner_b200 import configs, capi
from oracle import oracle_py as orc
cid=int(sys.argv[1]); B=int(sys.argv[2]); mu=float(sys.argv[3]) if len(sys.argv)>3 else 0.1
cfg = configs.config_for(cid, tol=1e-8); cfg.mu_init=mu
data = configs.generate(cid, B) if cid!=1 else configs.g1_instance()
t=time.time()
out = orc.step_batch(cfg, data, n_threads=8)
el=time.time()-t
st=out['status']; it=out['iters']
print("cfg",cid,"mu",mu,"conv %d/%d"%((st==0).sum(),B),"maxit",(st==1).sum(),"numerr",(st==2).sum(),"iters mean %.1f med %d max %d"%(it.mean(),np.median(it),it.max()),"time %.2fs"%el)
print("fail idx", np.nonzero(st!=0)[0][:20])
