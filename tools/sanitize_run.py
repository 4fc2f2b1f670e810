"""Small cold solves of every code path (batch in both solve modes, queue, line / moving obstacles, minimum time)
for compute-sanitizer:  compute-sanitizer --tool memcheck python tools/sanitize_run.py"""
import sys; sys.path.insert(0, '.')
import numpy as np
from mpc_local_planner_b200 import capi, configs
for cid, B in ((2, 48), (3, 24), (4, 32)):
    cfg = configs.config_for(cid, tol=1e-6)
    cfg.max_iter = 30
    data = configs.generate(cid, B)
    for mode in (capi.SOLVE_FUSED, capi.SOLVE_PHASED):
        s = capi.BatchSolver(cfg, B)
        s.set_option(capi.OPT_SOLVE_MODE, mode)
        out = s.step(data["x0"], data["xf"], data["u_prev"], data["u_prev_dt"], data["obstacles"], data["viapoints"])
        s.close()
    s = capi.BatchSolver(cfg, 16)
    out = s.solve_stream(data["x0"], data["xf"], data["u_prev"], data["u_prev_dt"], data["obstacles"], data["viapoints"])
    s.close()
    print("cfg", cid, "ok", int((out["status"] == 0).sum()), "converged of", B, flush=True)
cfg = configs.cfg2_integral_form(tol=1e-6); cfg.max_iter = 30; cfg.enable_dynamic_obstacles = 1
data = configs.with_line_obstacles(configs.with_moving_obstacles(configs.generate(2, 32)))
s = capi.BatchSolver(cfg, 32)
out = s.step(data["x0"], data["xf"], data["u_prev"], data["u_prev_dt"], data["obstacles"], None)
s.close()
print("line + moving obstacles ok", int((out["status"] == 0).sum()), flush=True)
# phase alignment with several CTAs per SM, forced on (few iterations: the gates are what is being checked)
cfg = configs.cfg2(tol=1e-6); cfg.max_iter = 6
data = configs.generate(2, 448)
s = capi.BatchSolver(cfg, 448)
s.set_option(capi.OPT_SM_PHASE_SYNC, 1)
out = s.step(data["x0"], data["xf"], data["u_prev"], data["u_prev_dt"], data["obstacles"], None)
s.set_option(capi.OPT_SM_PHASE_SYNC, 2)
out = s.step(data["x0"], data["xf"], data["u_prev"], data["u_prev_dt"], data["obstacles"], None)
s.close()
print("phase alignment ok", flush=True)
# costmaps: extraction, long obstacle lists through the association, the chained call, the feasibility check
rng = np.random.default_rng(3)
B, W, H, res = 6, 96, 80, 0.05
cost = rng.choice(np.array([0, 1, 100, 253], dtype=np.uint8), size=(B, H, W)); cost[rng.random((B, H, W)) < 0.02] = 254
origin = rng.uniform(-1, 1, (B, 2))
pose = np.concatenate([origin + rng.uniform(1.0, 2.0, (B, 2)), rng.uniform(-3, 3, (B, 1))], axis=1)
goal = pose.copy(); goal[:, 0] += 2.0 * np.cos(pose[:, 2]); goal[:, 1] += 2.0 * np.sin(pose[:, 2])
cfg = configs.cfg2(tol=1e-6); cfg.max_iter = 8
s = capi.BatchSolver(cfg, B)
(count, typ, par), found = s.costmap_obstacles(cost, origin, res, pose, 0.3, 256)
out = s.step(pose, goal, None, 0.2, (count, typ, par), None)
out = s.step_from_costmaps(pose, goal, cost, origin, res, 0.3, 256)
fp = np.array([[0.2, 0.15], [-0.2, 0.15], [-0.2, -0.15], [0.2, -0.15]])
ok = s.check_feasible(cost, origin, res, fp, 0.15, 0.2, look_ahead_idx=10)
s.close()
print("costmap paths ok, obstacles per robot", count.tolist(), "feasible", ok.astype(int).tolist(), flush=True)
