import sys; sys.path.insert(0,'.')
import numpy as np
from mpc_local_planner_b200 import configs, capi
from tests.emu import emu_py as emu
cid=int(sys.argv[1]); B=int(sys.argv[2])
cfg=configs.config_for(cid,tol=1e-6); data=configs.generate(cid,B)
res=[]
for b in range(B):
    e=emu.instance_from_batch(cfg,data,b)
    e.init(); e.associate()
    al=[]; dl=[]; st=-1
    for it in range(cfg.max_iter+1):
        if e.eval(): break
        if it==cfg.max_iter: break
        if e.kkt(): break
        e.linesearch()
        sc=e.field(capi.F_SCAL); al.append(sc[capi.SC_ALPHA]); dl.append(sc[capi.SC_DELTA])
    st=int(e.field(capi.F_SCAL)[capi.SC_STATUS]); 
    res.append((st,np.array(al),np.array(dl)))
conv=[r for r in res if r[0]==0]; fail=[r for r in res if r[0]!=0]
print("converged",len(conv),"failed",len(fail))
print("converged: min alpha per instance quantiles", np.quantile([r[1].min() for r in conv],[0,0.05,0.25,0.5]))
print("converged: max delta per instance quantiles", np.quantile([r[2].max() for r in conv],[0.5,0.9,0.99,1]))
def runlen(a,thr):
    best=0;cur=0
    for v in a:
        cur=cur+1 if v<thr else 0; best=max(best,cur)
    return best
for thr in (1e-2,1e-3,1e-4):
    print("thr",thr,"converged: longest run of alpha<thr: max",max(runlen(r[1],thr) for r in conv), " failed: first iter where run>=3:", sorted([next((i for i in range(len(r[1])) if i>=2 and (r[1][i-2:i+1]<thr).all()),999) for r in fail])[:40])
print("failed final iters", sorted(len(r[1]) for r in fail))
