import sys; sys.path.insert(0,'.')
import numpy as np
from mpc_local_planner_b200 import configs, capi
from oracle import oracle_py as orc
b=int(sys.argv[1])
cfg=configs.cfg2(tol=1e-8); data=configs.generate(2,b+1)
print("x0",data['x0'][b],"xf",data['xf'][b])
print("obst\n",data['obstacles'][2][b][:,[0,1,4]])
inst=orc.instance_from_batch(cfg,data,b)
inst.init_cold(); X0=inst.arr('X').copy(); inst.associate(); 
print("obsidx\n",inst.arr('OBSIDX').astype(int))
u,x,r=inst.step()
print(r.status,r.iters)
P=inst.arr('X')
for k in range(30,45): print(k, X0[:,k].round(3), P[:,k].round(3), inst.arr('U')[:,k].round(3), inst.arr('G')[8:,k].round(4))
