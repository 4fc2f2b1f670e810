import sys; sys.path.insert(0,'.')
import numpy as np, time
from mpc_local_planner_b200 import configs, capi
from oracle import oracle_py as orc
B=32
for name,mod in [("noobs", lambda c,d: d['obstacles'][0].__setitem__(slice(None),0)),
                 ("noobs_norate", lambda c,d: (d['obstacles'][0].__setitem__(slice(None),0), c.du_lb.__setitem__(slice(None),[-1e30,-1e30]), c.du_ub.__setitem__(slice(None),[1e30,1e30]))),
                 ("obs_norate", lambda c,d: (c.du_lb.__setitem__(slice(None),[-1e30,-1e30]), c.du_ub.__setitem__(slice(None),[1e30,1e30]))),
                 ("full", lambda c,d: None)]:
    cfg = configs.cfg2(tol=1e-8); data = configs.generate(2,B)
    mod(cfg,data)
    out = orc.step_batch(cfg, data, n_threads=8)
    st=out['status']; it=out['iters']
    print(name,"conv %d/%d"%((st==0).sum(),B),"maxit",(st==1).sum(),"numerr",(st==2).sum(),"iters mean %.1f med %d max %d"%(it.mean(),np.median(it),it.max()))
