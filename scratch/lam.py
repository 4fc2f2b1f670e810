import sys; sys.path.insert(0,'.')
import numpy as np
from mpc_local_planner_b200 import configs, capi
from oracle import oracle_py as orc
for cid in (2,3):
  cfg = configs.config_for(cid, tol=1e-8); data = configs.generate(cid,16)
  for b in range(16):
    inst = orc.instance_from_batch(cfg, data, b)
    u,x,r = inst.step()
    if r.status==0:
        lam=inst.arr('LAM'); 
        print(cid,b,"iters",r.iters,"max lam obs %.3g"%lam[8:].max(), "max lam rate %.3g"%lam[4:8].max(), "max lam bnd %.3g"%lam[:4].max(), "max nu %.3g"%np.abs(inst.arr('NU')).max())
