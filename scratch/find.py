import sys; sys.path.insert(0,'.')
import numpy as np
from mpc_local_planner_b200 import configs, capi
from oracle import oracle_py as orc
cid=int(sys.argv[1]); B=int(sys.argv[2])
cfg = configs.config_for(cid, tol=1e-8)
data = configs.generate(cid, B)
out = orc.step_batch(cfg, data, n_threads=8)
print("numerr", np.nonzero(out['status']==2)[0][:10], "maxit", np.nonzero(out['status']==1)[0][:10])
print(out['iters'][out['status']==2][:10])
