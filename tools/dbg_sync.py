import sys; sys.path.insert(0, '.')
import numpy as np, time
print("start", flush=True)
from mpc_local_planner_b200 import capi, configs
import os
if os.environ.get('LIBP'): capi.load_library(os.environ['LIBP']); capi._LIB = capi.load_library(os.environ['LIBP'])
cfg = configs.config_for(2); B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
data = configs.generate(2, B)
sync = int(sys.argv[1])
s = capi.BatchSolver(cfg, B); print("created", flush=True)

if sync >= 0: s.set_option(capi.OPT_SM_PHASE_SYNC, sync); print("opt", flush=True)
s.upload(data["x0"], data["xf"], data["u_prev"], data["u_prev_dt"], data["obstacles"], data["viapoints"]); print("uploaded", flush=True)
t = s.solve_resident(cold=True); print("solved", t, flush=True)
out = s.fetch(); print("status", np.bincount(out["status"]), s.stats()["gate_ms"], flush=True)
