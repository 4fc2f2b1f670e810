"""One planning cycle from the costmaps for a fleet: 1024 robots x 400 x 400 cells, ~160 LETHAL cells in reach of each
robot -> mpcb200_step_batch_costmap: maps H2D, extraction, association over the raw lists in global memory, solve; lists stay on
the device.  Prints one JSON line.  Usage: python tools/costmap_chain_bench.py [robots] [size] [max_per_instance]"""
import sys, time, json; sys.path.insert(0, '.')
import numpy as np
from mpc_local_planner_b200 import capi, configs
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
W = H = int(sys.argv[2]) if len(sys.argv) > 2 else 400
M = int(sys.argv[3]) if len(sys.argv) > 3 else 512
rng = np.random.default_rng(0)
res = 0.05
cost = rng.choice(np.array([0, 1, 100, 253], dtype=np.uint8), size=(B, H, W))
cost[rng.random((B, H, W)) < 0.002] = 254
origin = rng.uniform(-3, 3, (B, 2))
pose = np.concatenate([origin + 0.5 * res * np.array([W, H]) + rng.uniform(-2, 2, (B, 2)), rng.uniform(-np.pi, np.pi, (B, 1))], axis=1)
goal = pose.copy(); goal[:, 0] += 4.0 * np.cos(pose[:, 2]); goal[:, 1] += 4.0 * np.sin(pose[:, 2])
yy, xx = np.mgrid[0:H, 0:W]
for b in range(B):   # free cells around start and goal
    cx, cy = origin[b, 0] + (xx + 0.5) * res, origin[b, 1] + (yy + 0.5) * res
    for p in (pose[b], goal[b]):
        cost[b][(cx - p[0]) ** 2 + (cy - p[1]) ** 2 < 0.6 ** 2] = 0
cfg = configs.cfg2(tol=1e-6)
s = capi.BatchSolver(cfg, B, device=0)
out = s.step_from_costmaps(pose, goal, cost, origin, res, 0.3, M)   # warm-up (allocations)
walls, solves, cms = [], [], []
for _ in range(5):
    s.reset()
    t = time.time()
    out = s.step_from_costmaps(pose, goal, cost, origin, res, 0.3, M)
    walls.append(time.time() - t); solves.append(out["solve_time_s"]); cms.append(s.costmap_last_ms())
SC = s.ws_read(capi.F_SCAL)
G = s.ws_read(capi.F_OBSGIDX)
print(json.dumps(dict(robots=B, size_x=W, size_y=H, max_per_instance=M, map_bytes=int(cost.nbytes), obstacles_found_mean=float(out["obst_found"].mean()),
                      obstacles_found_max=int(out["obst_found"].max()), resident_mean=float((G >= 0).sum(1).mean()), dropped_rows=int(SC[:, capi.SC_OBST_DROPPED].sum()),
                      converged_fraction=float((out["status"] == 0).mean()), wall_ms=float(np.median(walls) * 1e3), costmap_kernels_ms=float(np.median(cms)),
                      solve_kernel_ms=float(np.median(solves) * 1e3), mean_iterations=float(out["iters"].mean()))), flush=True)
# the same cycle with the lists going through the host (extraction call + step call)
walls2 = []
for _ in range(3):
    s.reset()
    t = time.time()
    (count, typ, par), found = s.costmap_obstacles(cost, origin, res, pose, 0.3, M)
    out2 = s.step(pose, goal, None, 0.0, (count, typ, par), None)
    walls2.append(time.time() - t)
assert (out2["status"] == out["status"]).all() and (out2["u_seq"] == out["u_seq"]).all()
print(json.dumps(dict(two_calls_through_the_host_wall_ms=float(np.median(walls2) * 1e3))))
s.close()
