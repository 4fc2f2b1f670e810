"""Whole-solve timings of the BASELINE configurations on one GPU: wall / device ms, converged solves/s, per-phase device
times (one extra solve with every phase bracketed) and the eval / KKT kernels timed alone.  Usage: python tools/solve_timings.py"""
import sys, time, json; sys.path.insert(0,'.')
import numpy as np
from mpc_local_planner_b200 import capi, configs
from oracle import oracle_py as orc
for cid,B,tol in ((2,1024,1e-6),(2,4096,1e-6),(3,1024,1e-6),(4,2048,1e-6)):
    cfg=configs.config_for(cid,tol=tol); data=configs.generate(cid,B)
    s=capi.BatchSolver(cfg,B)
    for rep in range(3):
        s.reset(); s.stats_reset()
        t=time.time(); out=s.step(data["x0"],data["xf"],data["u_prev"],data["u_prev_dt"],data["obstacles"],data["viapoints"]); el=time.time()-t
    conv=(out['status']==0).sum()
    s.set_timing(0x1f); s.reset(); s.stats_reset()
    out2=s.step(data["x0"],data["xf"],data["u_prev"],data["u_prev_dt"],data["obstacles"],data["viapoints"])
    st=s.stats(); s.set_timing(1<<3)
    assert (out2['status']==out['status']).all()
    print(json.dumps(dict(cfg=cid,B=B,conv=int(conv),wall_ms=el*1e3,dev_ms=out['solve_time_s']*1e3,solves_per_s=conv/el,stats=st,iters_mean=float(out['iters'].mean()))))
    # kernel timing
    s.reset(); s.upload(data["x0"],data["xf"],data["u_prev"],data["u_prev_dt"],data["obstacles"],data["viapoints"])
    s.run_phase(capi.PHASE_INIT); s.run_phase(capi.PHASE_ASSOCIATE); s.run_phase(capi.PHASE_EVAL)
    for ph,name in ((capi.PHASE_EVAL,'eval'),(capi.PHASE_KKT,'kkt')):
        ms=s.time_phase(ph,reps=10,flush_l2=True); ms2=s.time_phase(ph,reps=10,flush_l2=False)
        N=cfg.n; words=46+4*(cfg.du_ub[0]<1e29)+3*cfg.variable_dt
        byts=B*8*(words*(N-1)+12)
        print("   ",name,"ms flushed %.4f warm %.4f"%(ms,ms2), "alg GB/s (kkt) %.1f"%(byts/ms/1e6))
    s.close()
