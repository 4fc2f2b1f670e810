import sys; sys.path.insert(0,'.')
import numpy as np
from mpc_local_planner_b200 import configs, capi
from tests.emu import emu_py as emu
cid=2; B=96
cfg=configs.config_for(cid,tol=1e-6); data=configs.generate(cid,B)
trans={(0,0):0,(0,1):0,(1,0):0,(1,1):0}; tries=[]; 
sweeps_per_iter=np.zeros(101); active_per_iter=np.zeros(101); maxsweeps=np.zeros((B,101))
for b in range(B):
    e=emu.instance_from_batch(cfg,data,b); e.init(); e.associate()
    prev=0
    for it in range(cfg.max_iter+1):
        if e.eval(): break
        if it==cfg.max_iter: break
        n0=e.field(capi.F_SCAL)[capi.SC_NREG]
        if e.kkt(): break
        nreg=int(e.field(capi.F_SCAL)[capi.SC_NREG]-n0)
        cur=1 if e.field(capi.F_SCAL)[capi.SC_DELTA]>0 else 0
        trans[(prev,cur)]+=1; prev=cur; tries.append(nreg+1)
        sweeps_per_iter[it]+=nreg+1; active_per_iter[it]+=1; maxsweeps[b,it]=nreg+1
        e.linesearch()
print("transitions (prev reg?, cur reg?)",trans)
tries=np.array(tries); print("sweeps per kkt call: mean %.2f, hist"%tries.mean(), np.bincount(tries)[:14])
# emulate tile max: groups of 32 instances
for g in range(B//32):
    m=maxsweeps[g*32:(g+1)*32].max(axis=0); print("tile",g,"sum over iters of max sweeps",m.sum(),"vs mean-lane", maxsweeps[g*32:(g+1)*32].sum()/32)
