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
