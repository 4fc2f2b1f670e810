import sys; sys.path.insert(0,'.')
import numpy as np
from mpc_local_planner_b200 import configs, capi
from oracle import oracle_py as orc
from tests.emu import emu_py as emu
def cmp(name,a,b):
    d=np.abs(a-b).max(); s=max(np.abs(b).max(),1e-300)
    print("  %-8s maxabs %.3e rel %.3e"%(name,d,d/s))
for cid in (1,2,3,4):
    cfg=configs.config_for(cid,tol=1e-8); data=configs.generate(cid,3) if cid!=1 else configs.g1_instance()
    b=0 if cid==1 else 2
    o=orc.instance_from_batch(cfg,data,b); e=emu.instance_from_batch(cfg,data,b)
    o.init_cold(); o.associate(); o.L.orc_project_init(orc.C.byref(o.p),o.ws); o.L.orc_init_controls(orc.C.byref(o.p),o.ws); o.init_duals()
    e.init(); e.associate()
    print("cfg",cid)
    cmp("X",e.field(capi.F_X),o.arr('X')); cmp("U",e.field(capi.F_U),o.arr('U')); cmp("S",e.field(capi.F_S),o.arr('S')); cmp("LAM",e.field(capi.F_LAM),o.arr('LAM')); cmp("OBS",e.field(capi.F_OBSIDX),o.arr('OBSIDX'))
    for it in range(3):
        o.eval(); e.eval()
        cmp("KKT",e.field(capi.F_KKT),o.arr('KKT')); cmp("SCAL",e.field(capi.F_SCAL)[[0,1,4,5,7,8,12,13,16,17]],o.arr('SCAL')[[0,1,4,5,7,8,12,13,16,17]])
        ro=o.kkt_solve(0.0); re=e.kkt()
        print("  kkt rc",ro,re, "delta", e.field(capi.F_SCAL)[capi.SC_DELTA])
        if ro==0 and e.field(capi.F_SCAL)[capi.SC_DELTA]==0:
            cmp("STEP",e.field(capi.F_STEP),o.arr('STEP')); cmp("DDT",e.field(capi.F_SCAL)[[capi.SC_DDT]],o.arr('SCAL')[[capi.SC_DDT]])
        # advance both with emulator's line search result copied into oracle
        e.linesearch()
        o.arr('X')[:]=e.field(capi.F_X); o.arr('U')[:]=e.field(capi.F_U); o.arr('NU')[:]=e.field(capi.F_NU); o.arr('S')[:]=e.field(capi.F_S); o.arr('LAM')[:]=e.field(capi.F_LAM)
        o.arr('SCAL')[capi.SC_DT]=e.field(capi.F_SCAL)[capi.SC_DT]; o.arr('SCAL')[capi.SC_MU]=e.field(capi.F_SCAL)[capi.SC_MU]
