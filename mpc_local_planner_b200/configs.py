"""BASELINE.json configurations restated as concrete solver configs + synthetic instance generators.

All parameter values are taken from the reference's example YAML files (cited inline); the random instance
distributions are the ones fixed in SURVEY.md 8(d).  RNG: numpy PCG64, seed 0xB2000000 + 1000*config_id, one
spawned child stream per instance index, so instance i is identical for every batch size and GPU count.
"""
import math

import numpy as np

from . import capi

SEED_BASE = 0xB2000000

# EX/cfg/carlike/mpc_local_planner_params.yaml:28
CARLIKE_POLYGON = [[0.25, -0.05], [0.18, -0.05], [0.18, -0.18], [-0.19, -0.18], [-0.25, 0.0], [-0.19, 0.18],
                   [0.18, 0.18], [0.18, 0.05], [0.25, 0.05]]


def _diag(c_arr, vals):
    n = len(vals)
    for i in range(n * n):
        c_arr[i] = 0.0
    for i, v in enumerate(vals):
        c_arr[i * n + i] = v


def cfg1(tol=1e-6):
    """single unicycle minimum-time OCP, N=20 (P/cfg/test_mpc_optim_node.yaml:5-10,32,35,44-53,64)."""
    c = capi.default_config()
    c.robot_type = capi.ROBOT_UNICYCLE
    c.u_lb[:] = [-0.2, -0.3]
    c.u_ub[:] = [0.4, 0.3]
    c.n, c.dt_ref = 20, 0.3
    c.variable_dt, c.dt_lb, c.dt_ub = 1, 0.0, 10.0
    c.xf_fixed[:] = [1, 1, 1]
    c.objective = capi.OBJ_MINIMUM_TIME
    c.min_obstacle_dist, c.force_inclusion_dist, c.cutoff_dist = 0.5, 0.5, 2.0
    c.footprint_type = capi.FOOTPRINT_POINT
    c.k_max_obstacles_per_stage = 3
    c.tol = tol
    return c


def cfg2(n=50, tol=1e-6):
    """unicycle quadratic-form, N=50, fixed dt
    (EX/cfg/diff_drive/mpc_local_planner_params_quadratic_form.yaml:9-14,18,23-26,34-41,53-61)."""
    c = capi.default_config()
    c.robot_type = capi.ROBOT_UNICYCLE
    c.u_lb[:] = [-0.2, -0.3]
    c.u_ub[:] = [0.4, 0.3]
    c.du_lb[:] = [-0.2, -0.2]
    c.du_ub[:] = [0.2, 0.2]
    c.n, c.dt_ref = n, 0.3
    c.variable_dt = 0
    c.xf_fixed[:] = [0, 0, 0]
    c.objective = capi.OBJ_QUADRATIC_FORM
    _diag(c.Q, [2.0, 2.0, 0.25])
    _diag(c.R, [0.1, 0.05])
    c.terminal_cost = 1
    _diag(c.Qf, [10.0, 10.0, 0.5])
    c.min_obstacle_dist, c.force_inclusion_dist, c.cutoff_dist = 0.2, 0.5, 2.5
    c.footprint_type = capi.FOOTPRINT_POINT
    c.k_max_obstacles_per_stage = 5
    c.tol = tol
    return c


def cfg3(n=80, tol=1e-6):
    """carlike (simple_car rear drive, L=0.4) minimum-time, polygon footprint
    (EX/cfg/carlike/mpc_local_planner_params.yaml:8-16,28,34-37,46-55,65)."""
    c = capi.default_config()
    c.robot_type = capi.ROBOT_SIMPLE_CAR
    c.wheelbase = 0.4
    c.u_lb[:] = [-0.2, -1.4]
    c.u_ub[:] = [0.4, 1.4]
    c.du_lb[:] = [-0.5, -0.5]
    c.du_ub[:] = [0.5, 0.5]
    c.n, c.dt_ref = n, 0.3
    c.variable_dt, c.dt_lb, c.dt_ub = 1, 0.0, 10.0
    c.xf_fixed[:] = [1, 1, 1]
    c.objective = capi.OBJ_MINIMUM_TIME
    c.min_obstacle_dist, c.force_inclusion_dist, c.cutoff_dist = 0.27, 0.5, 2.5
    c.footprint_type = capi.FOOTPRINT_POLYGON
    c.n_poly = len(CARLIKE_POLYGON)
    for i, (x, y) in enumerate(CARLIKE_POLYGON):
        c.poly_xy[2 * i], c.poly_xy[2 * i + 1] = x, y
    c.k_max_obstacles_per_stage = 5
    c.tol = tol
    return c


def cfg4(n=50, tol=1e-6):
    """cfg2 + via-point attraction (SURVEY 8d reading A; weight 8.0 from P/cfg/test_mpc_optim_node.yaml:71)."""
    c = cfg2(n, tol)
    c.vp_attraction_with_quadratic = 1
    c.vp_position_weight = 8.0
    c.vp_orientation_weight = 0.0
    c.vp_ordered = 0
    return c


def cfg5(n, tol=1e-6):
    return cfg2(n, tol)


def _streams(config_id, B, first=0):
    ss = np.random.SeedSequence(SEED_BASE + 1000 * config_id)
    children = ss.spawn(first + B)[first:]
    return [np.random.Generator(np.random.PCG64(s)) for s in children]


def _gen_common(rng, r_lo, r_hi, d_min, n_obst, circles_only=True):
    th0 = rng.uniform(-math.pi / 4, math.pi / 4)
    r = rng.uniform(r_lo, r_hi)
    bearing = rng.uniform(-math.pi / 4, math.pi / 4)
    gh = rng.uniform(-math.pi / 2, math.pi / 2)
    x0 = np.array([0.0, 0.0, th0])
    xf = np.array([r * math.cos(bearing), r * math.sin(bearing), gh])
    types = np.zeros(n_obst, dtype=np.int32)
    params = np.zeros((n_obst, capi.OBST_STRIDE))
    dirv = xf[:2] / r
    nrm = np.array([-dirv[1], dirv[0]])
    for j in range(n_obst):
        while True:
            rad = rng.uniform(0.1, 0.3)
            s = rng.uniform(0.2, 0.8)
            lat = rng.uniform(-1.0, 1.0)
            is_circle = True if circles_only else bool(rng.integers(0, 2))
            ctr = s * xf[:2] + lat * nrm
            rr = rad if is_circle else 0.0
            clear = rr + d_min + 0.3
            if np.linalg.norm(ctr - x0[:2]) > clear and np.linalg.norm(ctr - xf[:2]) > clear:
                break
        types[j] = capi.OBST_CIRCLE if is_circle else capi.OBST_POINT
        params[j, 0:2] = ctr
        params[j, 4] = rr
    return x0, xf, types, params


def generate(config_id, B, first=0, n=None):
    """Synthetic instances of BASELINE config `config_id` (2..5). Returns dict of numpy arrays ready for
    BatchSolver.step / the oracle: x0, xf, u_prev, u_prev_dt, obstacles=(count,types,params), viapoints or None."""
    rngs = _streams(config_id, B, first)
    n_obst = 5
    x0 = np.zeros((B, 3))
    xf = np.zeros((B, 3))
    count = np.full(B, n_obst, dtype=np.int32)
    types = np.zeros((B, n_obst), dtype=np.int32)
    params = np.zeros((B, n_obst, capi.OBST_STRIDE))
    vps = None
    if config_id == 4:
        vps = (np.full(B, 2, dtype=np.int32), np.zeros((B, 2, 3)))
    for b, rng in enumerate(rngs):
        if config_id in (2, 4):
            a = _gen_common(rng, 3.0, 6.0, 0.2, n_obst)
        elif config_id == 3:
            a = _gen_common(rng, 4.0, 8.0, 0.27, n_obst, circles_only=False)
        elif config_id == 5:
            nn = n or 50
            a = _gen_common(rng, 0.06 * nn, 0.12 * nn, 0.2, n_obst)
        else:
            raise ValueError("generate() supports config ids 2..5")
        x0[b], xf[b], types[b], params[b] = a
        if config_id == 4:
            r = np.linalg.norm(xf[b, :2])
            dirv = xf[b, :2] / r
            nrm = np.array([-dirv[1], dirv[0]])
            for j, frac in enumerate((1.0 / 3.0, 2.0 / 3.0)):
                lat = rng.uniform(-0.5, 0.5)
                vps[1][b, j, :2] = frac * xf[b, :2] + lat * nrm
                vps[1][b, j, 2] = 0.0
    return dict(x0=x0, xf=xf, u_prev=np.zeros((B, 2)), u_prev_dt=0.2, obstacles=(count, types, params), viapoints=vps)


def cfg2_integral_form(n=50, tol=1e-6):
    """cfg 2 with `planning/objective/quadratic_form/integral_form: true` (left sum) and a free dt in [0.05, 1.0] s --
    not a BASELINE configuration; exercises the dt-dependence of the integrated running cost."""
    c = cfg2(n, tol)
    c.quadratic_integral_form = 1
    c.variable_dt, c.dt_lb, c.dt_ub = 1, 0.05, 1.0
    return c


def cfg2_trapezoidal(n=50, tol=1e-6, variable_dt=True):
    """cfg 2 in integral form integrated by the trapezoidal rule (`grid/cost_integration_method: trapezoidal_rule`), with a
    free dt in [0.05, 1.0] s (the end term dt/2 l(x_{N-1}) then couples the free final state and dt) or the fixed dt of
    cfg 2 -- not a BASELINE configuration."""
    c = cfg2_integral_form(n, tol) if variable_dt else cfg2(n, tol)
    c.quadratic_integral_form = 1
    c.cost_integration = capi.COST_TRAPEZOIDAL
    return c


def cfg2_hybrid_min_time(n=50, tol=1e-6, integral_form=False):
    """cfg 2 as `quadratic_form/hybrid_cost_minimum_time: true` allows it (src/controller.cpp:595-620): zero state weights,
    control weights R, the minimum-time term, a free dt in [0.05, 1.0] s and a fixed final state (without state weights nothing
    else pulls the robot to the goal; the terminal cost edge disappears with it) -- not a BASELINE configuration."""
    c = cfg2(n, tol)
    for i in range(9):
        c.Q[i] = 0.0
    c.hybrid_cost_minimum_time = 1
    c.quadratic_integral_form = int(integral_form)
    c.variable_dt, c.dt_lb, c.dt_ub = 1, 0.05, 1.0
    c.xf_fixed[0] = c.xf_fixed[1] = c.xf_fixed[2] = 1
    return c


def cfg2_terminal_ball(n=50, tol=1e-6, gamma=0.05):
    """cfg 2 with `planning/terminal_constraint/type: l2_ball` (TerminalBallSE2: d' S d - gamma <= 0 on the final state),
    S = diag(1, 1, 0.5) -- not a BASELINE configuration."""
    c = cfg2(n, tol)
    c.terminal_ball = 1
    _diag(c.terminal_ball_S, [1.0, 1.0, 0.5])
    c.terminal_ball_gamma = gamma
    return c


def with_moving_obstacles(data, seed=0, vmax=0.15):
    """Variant of a batch in which every obstacle moves with a constant velocity ~U(-vmax, vmax)^2 (dynamic obstacles: set
    `enable_dynamic_obstacles` in the config to make the solver use the predicted positions)."""
    count, types, params = (a.copy() for a in data["obstacles"])
    rng = np.random.Generator(np.random.PCG64(np.random.SeedSequence(SEED_BASE + 99 + seed)))
    params[:, :, 5:7] = rng.uniform(-vmax, vmax, size=params[:, :, 5:7].shape)
    out = dict(data)
    out["obstacles"] = (count, types, params)
    return out


def with_line_obstacles(data, seed=0, every=2):
    """Variant of a generated batch in which every `every`-th obstacle is a LINE obstacle (a wall segment through the
    original centre, random direction, half-length radius + 0.25 m) -- exercises the LineObstacle distance of SURVEY App. B.3."""
    count, types, params = (a.copy() for a in data["obstacles"])
    rng = np.random.Generator(np.random.PCG64(np.random.SeedSequence(SEED_BASE + 77 + seed)))
    B, M = types.shape
    for b in range(B):
        for j in range(0, M, every):
            c = params[b, j, 0:2].copy()
            half = params[b, j, 4] + 0.25
            ang = rng.uniform(0.0, math.pi)
            d = half * np.array([math.cos(ang), math.sin(ang)])
            params[b, j, 0:2] = c - d
            params[b, j, 2:4] = c + d
            params[b, j, 4] = 0.0
            types[b, j] = capi.OBST_LINE
    out = dict(data)
    out["obstacles"] = (count, types, params)
    return out


def config_for(config_id, n=None, tol=1e-6):
    if config_id == 1:
        return cfg1(tol)
    if config_id == 2:
        return cfg2(n or 50, tol)
    if config_id == 3:
        return cfg3(n or 80, tol)
    if config_id == 4:
        return cfg4(n or 50, tol)
    if config_id == 5:
        return cfg5(n or 50, tol)
    raise ValueError(config_id)


def g1_instance():
    """Scenario G1: the reference's only fixed scenario (src/test_mpc_optim_node.cpp:67-69,105-106)."""
    x0 = np.array([[0.0, 0.0, 0.0]])
    xf = np.array([[5.0, 2.0, 0.0]])
    types = np.zeros((1, 3), dtype=np.int32)
    params = np.zeros((1, 3, capi.OBST_STRIDE))
    params[0, :, 0:2] = [[-3.0, 1.0], [6.0, 2.0], [4.0, 0.1]]
    return dict(x0=x0, xf=xf, u_prev=np.zeros((1, 2)), u_prev_dt=0.0,
                obstacles=(np.array([3], dtype=np.int32), types, params), viapoints=None)
