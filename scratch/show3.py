import sys; sys.path.insert(0,'.')
import numpy as np
from mpc_local_planner_b200 import configs, capi
from oracle import oracle_py as orc
cid=int(sys.argv[1]); b=int(sys.argv[2])
cfg=configs.config_for(cid,tol=1e-8); data=configs.generate(cid,b+1)
print("x0",data['x0'][b],"xf",data['xf'][b]); print(data['obstacles'][2][b][:,[0,1,4]], data['obstacles'][1][b])
inst=orc.instance_from_batch(cfg,data,b)
u,x,r=inst.step()
print(r.status,r.iters,r.dt)
np.set_printoptions(linewidth=200,precision=3,suppress=True)
print(np.hstack([x[::4],u[::4]]))
