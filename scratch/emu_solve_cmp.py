import sys; sys.path.insert(0,'.')
import numpy as np, time
from mpc_local_planner_b200 import configs, capi
from oracle import oracle_py as orc
from tests.emu import emu_py as emu
for cid,B in ((1,1),(2,24),(3,12),(4,12)):
    cfg=configs.config_for(cid,tol=1e-8); data=configs.generate(cid,B) if cid!=1 else configs.g1_instance()
    out=orc.step_batch(cfg,data,n_threads=8)
    agree=0; both=0; worst=0
    for b in range(B):
        e=emu.instance_from_batch(cfg,data,b)
        st=e.solve()
        u,x=e.outputs()
        it=int(e.field(capi.F_SCAL)[capi.SC_ITER])
        du=np.abs(u-out['u_seq'][b]).max()
        same = (st==out['status'][b])
        if st==0 and out['status'][b]==0:
            both+=1; worst=max(worst,du)
        agree+=same
        if not same or (st==0 and du>1e-6): print("  cfg",cid,"b",b,"emu st",st,"it",it,"orc st",out['status'][b],"it",out['iters'][b],"du %.2e"%du)
    print("cfg",cid,"status agree %d/%d"%(agree,B),"both converged",both,"worst du %.2e"%worst)
