import sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from mpc_local_planner_b200 import capi, configs
from test_gpu_variants import _costmap_scene
B, M = 4, 2048
cost, origin, res, pose, goal = _costmap_scene(B, 0.12, 12)
cfg = configs.cfg2(tol=1e-6)
s = capi.BatchSolver(cfg, B, device=0)
(count, typ, par), found = s.costmap_obstacles(cost, origin, res, pose, 0.3, M)
print("count", count, found)
s.upload(pose, goal, None, 0.2, (count, typ, par), None)
s.run_phase(capi.PHASE_INIT); s.run_phase(capi.PHASE_ASSOCIATE)
SC = s.ws_read(capi.F_SCAL); G = s.ws_read(capi.F_OBSGIDX); O = s.ws_read(capi.F_OBSIDX)
print("dropped", SC[:, capi.SC_OBST_DROPPED], "valid", SC[:, capi.SC_VALID], "res", (G >= 0).sum(1), "rows", (O >= 0).sum((1, 2)))
out = s.step(pose, goal, None, 0.2, (count, typ, par), None)
SC = s.ws_read(capi.F_SCAL); G = s.ws_read(capi.F_OBSGIDX); O = s.ws_read(capi.F_OBSIDX)
print("status", out["status"], "dropped", SC[:, capi.SC_OBST_DROPPED], "res", (G >= 0).sum(1), "rows", (O >= 0).sum((1, 2)))
