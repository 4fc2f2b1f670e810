import sys; sys.path.insert(0,'.')
import numpy as np
from mpc_local_planner_b200 import configs, capi
from oracle import oracle_py as orc
cfg=configs.cfg1(1e-8); data=configs.g1_instance()
inst=orc.instance_from_batch(cfg,data,0)
inst.init_cold(); inst.associate(); inst.init_duals(); inst.eval()
for d in (0,1e-4,1e-2,1,100):
    print(d, inst.kkt_solve(d))
K=inst.arr('KKT'); print("H diag stage 5", K[[0,5,9,12,14],5], "HN", K[[0,5,9],19], "htt", inst.arr('SCAL')[capi.SC_HTT])
print(inst.arr('OBSIDX'))
