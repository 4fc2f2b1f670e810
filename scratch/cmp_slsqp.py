import sys,pickle; sys.path.insert(0,'.')
import numpy as np
from mpc_local_planner_b200 import configs
from oracle import oracle_py as orc
res=pickle.load(open('scratch/slsqp_cfg2_pre1.pkl','rb'))
cfg=configs.config_for(2,tol=1e-8); data=configs.generate(2,64)
out=orc.step_batch(cfg,data,n_threads=4)
ok=[r['b'] for r in res if 'err' not in r and r['ce']<1e-8 and r['ci']>-1e-8]
oc=list(np.nonzero(out['status']==0)[0])
print("slsqp feasible",len(ok),"oracle conv",len(oc),"both",len(set(ok)&set(oc)),"only slsqp",sorted(set(ok)-set(oc)),"only oracle",sorted(set(oc)-set(ok)))
for r in res:
    b=r['b']
    if b in ok and b in oc:
        inst=orc.instance_from_batch(cfg,data,b); u,x,rr=inst.step()
        du=np.abs(r['U'].T-u[:-1]).max()
        print(b,"f slsqp %.6f oracle %.6f"%(r['f'],rr.objective),"du %.2e"%du, "nit",r['nit'],"it",rr.iters)
