"""Device time of cold batch solves in both solve modes and of the KKT kernel alone (cfg given on the command line).
usage: python tools/quick_bench.py [cfg=2] [B=1024] [reps=5]"""
import sys, time; sys.path.insert(0, '.')
import numpy as np
from mpc_local_planner_b200 import capi, configs
cid = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
cfg = configs.config_for(cid)
data = configs.generate(cid, B)
res = {}
for mode, name in ((capi.SOLVE_FUSED, "fused"), (capi.SOLVE_PHASED, "phased")):
    s = capi.BatchSolver(cfg, B)
    s.set_option(capi.OPT_SOLVE_MODE, mode)
    s.upload(data["x0"], data["xf"], data["u_prev"], data["u_prev_dt"], data["obstacles"], data["viapoints"])
    ts = []
    for r in range(reps):
        s.flush_l2(); s.stats_reset()
        ts.append(s.solve_resident(cold=True))
    out = s.fetch(); st = s.stats()
    res[name] = out
    conv = int((out["status"] == 0).sum())
    print(f"cfg {cid} B {B} {name:6s}: {min(ts)*1e3:8.3f} ms (median {np.median(ts)*1e3:.3f})  converged {conv}  "
          f"=> {conv/min(ts):,.0f} solves/s  iters mean {out['iters'].mean():.1f} max {out['iters'].max()}  "
          f"phase ms {[round(x,3) for x in st['ms']]} kkt_inst {st['kkt_instances']} sweeps {st['kkt_sweeps']}", flush=True)
    if name == "phased":
        s.upload(data["x0"], data["xf"], data["u_prev"], data["u_prev_dt"], data["obstacles"], data["viapoints"])
        s.reset()
        s.run_phase(capi.PHASE_INIT); s.run_phase(capi.PHASE_ASSOCIATE); s.run_phase(capi.PHASE_EVAL)
        for ph, nm in ((capi.PHASE_EVAL, "eval"), (capi.PHASE_KKT, "kkt"), (capi.PHASE_LINESEARCH, "linesearch")):
            ms = s.time_phase(ph, reps=10, flush_l2=True); ms2 = s.time_phase(ph, reps=10, flush_l2=False)
            print(f"   {nm:10s} alone: {ms*1e3:7.1f} us (L2 flushed)  {ms2*1e3:7.1f} us (warm)", flush=True)
    s.close()
a, b = res["fused"], res["phased"]
print("fused == phased:", bool((a["status"] == b["status"]).all() and (a["iters"] == b["iters"]).all() and (a["u_seq"] == b["u_seq"]).all()),
      "max |du|", float(np.abs(a["u_seq"] - b["u_seq"]).max()))
