import sys; sys.path.insert(0,'.')
import numpy as np
from mpc_local_planner_b200 import configs, capi
from tests.emu import emu_py as emu
cid=2; B=128
cfg=configs.config_for(cid,tol=1e-6); data=configs.generate(cid,B)
sw=np.zeros((B,101)); bt=np.zeros((B,101)); act=np.zeros((B,101))
final=np.zeros(B,int)
for b in range(B):
    e=emu.instance_from_batch(cfg,data,b); e.init(); e.associate()
    for it in range(cfg.max_iter+1):
        if e.eval(): break
        if it==cfg.max_iter: break
        sc=e.field(capi.F_SCAL); n0=sc[capi.SC_NREG]; b0=sc[capi.SC_NBT]
        if e.kkt(): break
        e.linesearch()
        sc=e.field(capi.F_SCAL)
        sw[b,it]=sc[capi.SC_NREG]-n0+1; bt[b,it]=sc[capi.SC_NBT]-b0+1; act[b,it]=1
    final[b]=int(e.field(capi.F_SCAL)[capi.SC_STATUS])
print("converged",(final==0).sum(),"of",B)
print("iter: active, max sweeps, max trials, mean sweeps(active), mean trials")
for it in list(range(0,40,3))+list(range(40,100,10)):
    a=act[:,it]>0
    if a.sum()==0: break
    print(it, int(a.sum()), int(sw[a,it].max()), int(bt[a,it].max()), "%.2f"%sw[a,it].mean(), "%.2f"%bt[a,it].mean())
print("sum over iters of max sweeps", sw.max(axis=0).sum(), "of max trials", bt.max(axis=0).sum())
conv=final==0
print("converged only: sum of max sweeps", sw[conv].max(axis=0).sum(), "max trials", bt[conv].max(axis=0).sum(), "last iter", int(act[conv].sum(axis=1).max()))
