import sys; sys.path.insert(0,'.')
import numpy as np
from mpc_local_planner_b200 import configs, capi
from oracle import oracle_py as orc
from tests.emu import emu_py as emu
cid=3; B=24
cfg=configs.config_for(cid,tol=1e-8); data=configs.generate(cid,B)
out=orc.step_batch(cfg,data,n_threads=8)
st=[];its=[]
for b in range(B):
    e=emu.instance_from_batch(cfg,data,b); s=e.solve(); st.append(s); its.append(int(e.field(capi.F_SCAL)[capi.SC_ITER]))
print("orc",out['status'], out['iters']); print("emu",np.array(st), np.array(its))
for b in range(B):
    if out['status'][b]==0 and st[b]==0:
        e=emu.instance_from_batch(cfg,data,b); e.solve(); u,x=e.outputs()
        print(b, "du %.2e"%np.abs(u-out['u_seq'][b]).max(), "dt emu %.9f orc %.9f"%(e.field(capi.F_SCAL)[capi.SC_DT], out['dt'][b]), "kkt", out['kkt_err'][b], e.field(capi.F_SCAL)[capi.SC_ERR0])
