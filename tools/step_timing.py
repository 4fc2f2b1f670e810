"""Where a bench step's time goes: device time of the solve alone vs the wall time of flush + solve + host bookkeeping."""
import sys, time; sys.path.insert(0, '.')
import numpy as np, torch
from mpc_local_planner_b200 import capi, configs
cfg = configs.config_for(2); B = 1024
data = configs.generate(2, B)
s = capi.BatchSolver(cfg, B)
use_stream = len(sys.argv) > 1 and sys.argv[1] == "stream"
if use_stream:
    st = torch.cuda.Stream(); s.set_stream(st.cuda_stream)
s.upload(data["x0"], data["xf"], data["u_prev"], data["u_prev_dt"], data["obstacles"], data["viapoints"])
for _ in range(3): s.flush_l2(); s.solve_resident(cold=True)
torch.cuda.synchronize()
dev = []; t0 = time.time()
for _ in range(10):
    s.flush_l2(); dev.append(s.solve_resident(cold=True))
torch.cuda.synchronize(); wall = (time.time() - t0) / 10
print("torch stream" if use_stream else "own stream", ": device ms per solve", [round(x * 1e3, 2) for x in dev], " wall ms per step %.2f" % (wall * 1e3))
