"""Costmap -> point obstacles (mpcb200_costmap_obstacles) for a batch of robots: device time of the three kernels and the
rate at which map bytes are consumed, next to the oracle's loops on the host.
Usage: python tools/costmap_bench.py [robots] [size_x] [size_y]"""
import sys, time, json; sys.path.insert(0, '.')
import numpy as np
from mpc_local_planner_b200 import capi, configs
from oracle import oracle_py as orc
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
W = int(sys.argv[2]) if len(sys.argv) > 2 else 400
H = int(sys.argv[3]) if len(sys.argv) > 3 else 400
rng = np.random.default_rng(0)
cost = rng.choice(np.array([0, 1, 100, 253], dtype=np.uint8), size=(B, H, W))
cost[rng.random((B, H, W)) < 0.002] = 254
origin = rng.uniform(-3, 3, (B, 2)); res = 0.05
pose = np.concatenate([origin + 0.5 * res * np.array([W, H]), rng.uniform(-np.pi, np.pi, (B, 1))], axis=1)
s = capi.BatchSolver(configs.cfg2(), 1)
for M in (64, 512):
    s.costmap_obstacles(cost, origin, res, pose, 0.3, M)
    ms = []
    for _ in range(5):
        t = time.time()
        (count, typ, par), found = s.costmap_obstacles(cost, origin, res, pose, 0.3, M)
        wall = time.time() - t
        ms.append(s.costmap_last_ms())
    dev = float(np.median(ms))
    print(json.dumps(dict(robots=B, size_x=W, size_y=H, max_per_instance=M, map_bytes=int(cost.nbytes), device_ms=dev,
                          map_GBps=cost.nbytes / dev * 1e-6, wall_ms_with_copies=wall * 1e3, found_mean=float(found.mean()), cut=int((found > M).sum()))), flush=True)
t = time.time()
n = min(B, 64)
for b in range(n):
    orc.costmap_obstacles(cost[b], origin[b], res, pose[b], 0.3, 512)
cpu = (time.time() - t) / n
print(json.dumps(dict(cpu_oracle_ms_per_robot=cpu * 1e3, cpu_map_GBps_one_core=W * H / cpu * 1e-9)))
s.close()
