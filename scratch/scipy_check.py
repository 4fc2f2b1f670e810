import sys, time; sys.path.insert(0,'.')
import numpy as np
from scipy.optimize import minimize
from mpc_local_planner_b200 import configs, capi
from oracle import oracle_py as orc

def scipy_solve(cfg, data, b, preprocess=True, method='SLSQP'):
    inst=orc.instance_from_batch(cfg,data,b)
    N=inst.N
    inst.init_cold()
    if preprocess:
        inst.L.orc_init_controls(orc.C.byref(inst.p), inst.ws)
    inst.associate()
    if preprocess:
        inst.L.orc_project_init(orc.C.byref(inst.p), inst.ws)
    inst.init_duals()
    X=inst.arr('X'); U=inst.arr('U'); SC=inst.arr('SCAL')
    # decision vector: x_1..x_{N-1} (free comps), u_0..u_{N-2}, dt
    idx=[]
    for k in range(1,N):
        for i in range(3):
            if k==N-1 and cfg.xf_fixed[i]: continue
            idx.append(('x',i,k))
    for k in range(N-1):
        for i in range(2): idx.append(('u',i,k))
    if cfg.variable_dt: idx.append(('dt',0,0))
    def setz(z):
        for v,(t,i,k) in zip(z,idx):
            if t=='x': X[i,k]=v
            elif t=='u': U[i,k]=v
            else: SC[capi.SC_DT]=v
    def getz():
        return np.array([X[i,k] if t=='x' else (U[i,k] if t=='u' else SC[capi.SC_DT]) for (t,i,k) in idx])
    act = inst.arr('LAM')>0
    def fun(z):
        setz(z); inst.eval(); return SC[capi.SC_OBJ]
    def ceq(z):
        setz(z); inst.eval(); return inst.arr('KKT')[capi.K_E:capi.K_E+3,:N-1].ravel().copy()
    def cin(z):
        setz(z); inst.eval(); return -(inst.arr('G')[act]).copy()
    z0=getz()
    t=time.time()
    if method=='SLSQP':
        res=minimize(fun,z0,method='SLSQP',constraints=[{'type':'eq','fun':ceq},{'type':'ineq','fun':cin}],options=dict(maxiter=500,ftol=1e-12))
    else:
        from scipy.optimize import NonlinearConstraint
        res=minimize(fun,z0,method='trust-constr',constraints=[NonlinearConstraint(ceq,0,0),NonlinearConstraint(cin,0,np.inf)],options=dict(maxiter=3000,gtol=1e-9,xtol=1e-12))
    setz(res.x)
    ce=np.abs(ceq(res.x)).max(); ci=min(cin(res.x).min(),0)
    return res, ce, ci, time.time()-t, U[:, :N-1].copy(), SC[capi.SC_DT]

if __name__=='__main__':
    cid=int(sys.argv[1]); B=int(sys.argv[2])
    cfg=configs.config_for(cid,tol=1e-8); data=configs.generate(cid,B) if cid!=1 else configs.g1_instance()
    out=orc.step_batch(cfg,data,n_threads=8)
    for b in range(B):
        res,ce,ci,tt,U,dt=scipy_solve(cfg,data,b)
        du=np.abs(U.T-out['u_seq'][b,:-1]).max()
        print(b,"scipy ok",res.success,res.status,"nit",res.nit,"f %.6f"%res.fun,"ceq %.1e cin %.1e"%(ce,ci),"%.1fs"%tt,"| oracle st",out['status'][b],"it",out['iters'][b],"du %.2e"%du, flush=True)
