import sys, time, pickle; sys.path.insert(0,'.'); sys.path.insert(0,'scratch')
import numpy as np
from multiprocessing import Pool
from mpc_local_planner_b200 import configs
from scipy_check import scipy_solve
def work(args):
    cid,b,pre=args
    cfg=configs.config_for(cid,tol=1e-8); data=configs.generate(cid,b+1)
    try:
        res,ce,ci,tt,U,dt=scipy_solve(cfg,data,b,preprocess=pre)
        return dict(cid=cid,b=b,f=res.fun,ce=ce,ci=ci,U=U,dt=dt,nit=res.nit,time=tt,status=res.status)
    except Exception as e:
        return dict(cid=cid,b=b,err=str(e))
if __name__=='__main__':
    cid=int(sys.argv[1]); B=int(sys.argv[2]); pre=int(sys.argv[3])
    with Pool(8) as p:
        out=p.map(work,[(cid,b,bool(pre)) for b in range(B)])
    pickle.dump(out,open(f'scratch/slsqp_cfg{cid}_pre{pre}.pkl','wb'))
    ok=[o for o in out if 'err' not in o and o['ce']<1e-8 and o['ci']>-1e-8]
    print("cfg",cid,"pre",pre,"feasible",len(ok),"/",B, "idx",[o['b'] for o in ok])
