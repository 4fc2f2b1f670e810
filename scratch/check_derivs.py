import sys; sys.path.insert(0,'.')
import numpy as np
from mpc_local_planner_b200 import configs, capi
from oracle import oracle_py as orc

def pack(inst):
    N=inst.N
    return np.concatenate([inst.arr('X').ravel(), inst.arr('U').ravel(), [inst.arr('SCAL')[capi.SC_DT]]])
def unpack(inst, z):
    N=inst.N
    inst.arr('X')[:] = z[:3*N].reshape(3,N); inst.arr('U')[:] = z[3*N:5*N].reshape(2,N); inst.arr('SCAL')[capi.SC_DT]=z[5*N]

def lagr(inst):
    inst.eval()
    S=inst.arr('SCAL'); K=inst.arr('KKT'); N=inst.N
    e=K[capi.K_E:capi.K_E+3,:N-1]
    L = S[capi.SC_OBJ] + (inst.arr('NU')[:,:N-1]*e).sum() + (inst.arr('LAM')*(inst.arr('G')+inst.arr('S'))*(inst.arr('LAM')>0)).sum()
    return L

def check(cfgid, b=0, seed=0):
    cfg = configs.config_for(cfgid, tol=1e-8)
    if cfgid==1:
        data = configs.g1_instance()
    else:
        data = configs.generate(cfgid, b+1)
    inst = orc.instance_from_batch(cfg, data, b if cfgid!=1 else 0)
    N=inst.N
    inst.init_cold(); 
    rng=np.random.default_rng(seed)
    inst.arr('X')[:,1:] += 0.05*rng.standard_normal((3,N-1))
    inst.arr('U')[:] = 0.1*rng.standard_normal((2,N)); inst.arr('U')[:,N-1]=0
    inst.associate(); inst.init_duals()
    inst.arr('NU')[:] = rng.standard_normal((3,N)); inst.arr('NU')[:,N-1]=0
    act = inst.arr('LAM')>0
    inst.arr('LAM')[:] = np.where(act, rng.uniform(0.5,2,act.shape), 0)
    inst.arr('SCAL')[capi.SC_MU]=0.1
    z0=pack(inst)
    inst.eval()
    GL=inst.arr('GL').copy(); gl_dt=inst.ws.contents.gl_dt
    KKT=inst.arr('KKT').copy()
    # FD gradient
    n=len(z0); g_fd=np.zeros(n); h=1e-6
    for i in range(n):
        zp=z0.copy(); zp[i]+=h; unpack(inst,zp); Lp=lagr(inst)
        zm=z0.copy(); zm[i]-=h; unpack(inst,zm); Lm=lagr(inst)
        g_fd[i]=(Lp-Lm)/(2*h)
    unpack(inst,z0)
    gx=g_fd[:3*N].reshape(3,N); gu=g_fd[3*N:5*N].reshape(2,N); gdt=g_fd[5*N]
    print("grad x err", np.abs(gx[:,1:]-GL[:3,1:]).max(), "u err", np.abs(gu[:,:N-1]-GL[3:5,:N-1]).max(), "dt err", abs(gdt-gl_dt), gdt, gl_dt)
    return inst
for c in (1,2,3,4):
    print("cfg",c); check(c)

def dense_check(cfgid, b=0, seed=0):
    inst = check(cfgid,b,seed)
    cfg=inst.cfg; N=inst.N
    z0=pack(inst)
    n=len(z0)
    # free variable index list
    free=[]
    for k in range(1,N):
        for i in range(3):
            if k==N-1 and cfg.xf_fixed[i]: continue
            free.append(i*N+k)
    for k in range(N-1):
        for i in range(2): free.append(3*N+i*N+k)
    if cfg.variable_dt: free.append(5*N)
    free=np.array(free)
    def grads(z):
        unpack(inst,z); inst.eval()
        GL=inst.arr('GL'); g=np.concatenate([GL[:3].ravel(), GL[3:5].ravel(), [inst.ws.contents.gl_dt]])
        K=inst.arr('KKT'); e=K[capi.K_E:capi.K_E+3,:N-1].copy()
        G=inst.arr('G').copy()
        return g, e.T.ravel(), G
    h=1e-6
    W=np.zeros((n,n)); m=3*(N-1); Jc=np.zeros((m,n))
    RS=inst.RS; Jg=np.zeros((RS*N,n))
    for i in range(n):
        zp=z0.copy(); zp[i]+=h; gp,ep,Gp=grads(zp)
        zm=z0.copy(); zm[i]-=h; gm,em,Gm=grads(zm)
        W[:,i]=(gp-gm)/(2*h); Jc[:,i]=(ep-em)/(2*h); Jg[:,i]=((Gp-Gm)/(2*h)).ravel()
    unpack(inst,z0); inst.eval()
    S=inst.arr('S').ravel(); LAM=inst.arr('LAM').ravel(); G=inst.arr('G').ravel()
    act=LAM>0
    sig=np.where(act, LAM/S, 0.0)
    mu=inst.arr('SCAL')[capi.SC_MU]
    Hc = W + Jg.T@np.diag(sig)@Jg
    # gradient of objective: GL - Jc^T nu - Jg^T lam
    gl,_,_=grads(z0)
    nu=inst.arr('NU')[:,:N-1].T.ravel()
    gJ = gl - Jc.T@nu - Jg.T@np.where(act,LAM,0)
    r = np.where(act, G+S, 0)
    gt = gJ + Jg.T@(np.where(act, mu/S,0) + sig*r)
    e = inst.arr('KKT')[capi.K_E:capi.K_E+3,:N-1].T.ravel()
    nf=len(free)
    Kmat=np.zeros((nf+m,nf+m)); Kmat[:nf,:nf]=Hc[np.ix_(free,free)]; Kmat[:nf,nf:]=Jc[:,free].T; Kmat[nf:,:nf]=Jc[:,free]
    rhs=np.concatenate([-gt[free], -e])
    for delta in (0.0, 1e-2):
        Kd=Kmat.copy(); Kd[:nf,:nf]+=delta*np.eye(nf)
        sol=np.linalg.solve(Kd,rhs)
        ev=np.linalg.eigvalsh(Kd)
        rc=inst.kkt_solve(delta)
        STEP=inst.arr('STEP'); ddt=inst.arr('SCAL')[capi.SC_DDT]
        dz=np.zeros(n); dz[:3*N]=STEP[:3].ravel(); dz[3*N:5*N]=STEP[3:5].ravel(); dz[5*N]=ddt
        nup=STEP[5:8,:N-1].T.ravel()
        print(" delta",delta,"rc",rc,"inertia(+,-)",(ev>0).sum(),(ev<0).sum(),"expected",nf,m,
              "dz err",np.abs(dz[free]-sol[:nf]).max(),"/",np.abs(sol[:nf]).max(),"nu err",np.abs(nup-sol[nf:]).max(),"/",np.abs(sol[nf:]).max())
for c in (1,2,3,4):
    print("dense cfg",c); dense_check(c)
