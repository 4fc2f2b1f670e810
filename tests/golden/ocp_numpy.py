"""Independent numpy restatement of the optimal control problem behind Controller::step -- TEST INFRASTRUCTURE.

Written from the reference's source lines (cited per function; R = /root/reference/mpc_local_planner), NOT from oracle/ and
not through it: nothing here imports, loads or calls liboracle.so.  It exists to pin the oracle (and through it the CUDA path):
  * scipy SLSQP on THESE functions produces the golden fixtures (tests/golden/make_golden.py),
  * the association and row / cost functions are compared value by value with the oracle's (tests/test_oracle_functions.py).
Only what the parameter sets of the fixtures need is covered in the distance functions (point / circle / line obstacles,
all five footprint models); everything is plain loops -- small problems only.
"""
import math

import numpy as np

INF = 1e30
PI = math.pi


# ---- R/include/mpc_local_planner/utils/math_utils.h:81-103 -------------------------------------------------------------------
def normalize_theta(theta):
    if -PI <= theta < PI:
        return theta
    multiplier = math.floor(theta / (2 * PI))
    theta = theta - multiplier * 2 * PI
    if theta >= PI:
        theta -= 2 * PI
    if theta < -PI:
        theta += 2 * PI
    return theta


def interpolate_angle(a1, a2, factor):
    return normalize_theta(a1 + factor * normalize_theta(a2 - a1))


# ---- R/include/mpc_local_planner/systems/*.h -------------------------------------------------------------------------------
def dynamics(cfg, x, u):
    th = x[2]
    rt = cfg.robot_type
    if rt == 0:    # unicycle_robot.h:59-68
        return np.array([u[0] * math.cos(th), u[0] * math.sin(th), u[1]])
    if rt == 1:    # simple_car.h:68-77 (rear wheel driving)
        return np.array([u[0] * math.cos(th), u[0] * math.sin(th), u[0] * math.tan(u[1]) / cfg.wheelbase])
    if rt == 2:    # simple_car.h:131-141 (front wheel driving)
        return np.array([u[0] * math.cos(th), u[0] * math.sin(th), u[0] * math.sin(u[1]) / cfg.wheelbase])
    # kinematic_bicycle_model.h:65-77
    lr, lf = cfg.length_rear, cfg.length_front
    beta = math.atan(lr / (lf + lr) * math.tan(u[1]))
    return np.array([u[0] * math.cos(th + beta), u[0] * math.sin(th + beta), u[0] * math.sin(beta) / lr])


# ---- R/include/mpc_local_planner/optimal_control/fd_collocation_se2.h:54-69 (forward), :91-108 (midpoint) ------------------
def defect(cfg, x1, u1, x2, dt):
    if cfg.collocation == 0:
        f = dynamics(cfg, x1, u1)
        return np.array([(x2[0] - x1[0]) / dt - f[0], (x2[1] - x1[1]) / dt - f[1], normalize_theta(x2[2] - x1[2]) / dt - f[2]])
    xm = np.array([0.5 * (x1[0] + x2[0]), 0.5 * (x1[1] + x2[1]), interpolate_angle(x1[2], x2[2], 0.5)])
    f = dynamics(cfg, xm, u1)
    return np.array([(x2[0] - x1[0]) / dt - f[0], (x2[1] - x1[1]) / dt - f[1], normalize_theta(x2[2] - x1[2]) / dt - f[2]])


# ---- teb_local_planner geometry (used at R/src/optimal_control/stage_inequality_se2.cpp:109,173) ----------
def _pt_seg(p, a, b):
    d = b - a
    sq = float(d @ d)
    t = 0.0 if sq <= 0 else min(1.0, max(0.0, float((p - a) @ d) / sq))
    return float(np.linalg.norm(p - (a + t * d)))


def _seg_intersect(p1, p2, p3, p4):
    d1, d2 = p2 - p1, p4 - p3
    den = d1[0] * d2[1] - d1[1] * d2[0]
    if abs(den) < 1e-14:
        return False
    r = p3 - p1
    t = (r[0] * d2[1] - r[1] * d2[0]) / den
    u = (r[0] * d1[1] - r[1] * d1[0]) / den
    return 0 <= t <= 1 and 0 <= u <= 1


def _seg_seg(a, b, c, d):
    if _seg_intersect(a, b, c, d):
        return 0.0
    return min(_pt_seg(a, c, d), _pt_seg(b, c, d), _pt_seg(c, a, b), _pt_seg(d, a, b))


def _footprint_features(cfg, pose):
    """world-frame features of the footprint: list of (a, b, radius) -- a point / circle is a degenerate segment"""
    p = np.array(pose[:2]); th = pose[2]
    R = np.array([[math.cos(th), -math.sin(th)], [math.sin(th), math.cos(th)]])
    ft = cfg.footprint_type
    fp = cfg.footprint_params
    if ft == 0:
        return [(p, p, 0.0)]
    if ft == 1:
        return [(p, p, fp[0])]
    if ft == 2:
        h = R[:, 0]
        return [(p + fp[0] * h, p + fp[0] * h, fp[1]), (p - fp[2] * h, p - fp[2] * h, fp[3])]
    if ft == 3:
        return [(p + R @ np.array([fp[0], fp[1]]), p + R @ np.array([fp[2], fp[3]]), 0.0)]
    n = cfg.n_poly
    v = [p + R @ np.array([cfg.poly_xy[2 * i], cfg.poly_xy[2 * i + 1]]) for i in range(n)]
    if n == 1:
        return [(v[0], v[0], 0.0)]
    if n == 2:
        return [(v[0], v[1], 0.0)]
    return [(v[i], v[(i + 1) % n], 0.0) for i in range(n)]


def footprint_distance(cfg, pose, otype, op, t=0.0):
    """RobotFootprintModel::calculateDistance(pose, obstacle); obstacle: 0 point (x, y), 1 circle (x, y, -, -, r), 2 line (x0, y0,
    x1, y1); velocity op[5:7] moves it by t * v first (estimateSpatioTemporalDistance, stage_inequality_se2.cpp:177-189)."""
    v = np.array([op[5], op[6]]) if (cfg.enable_dynamic_obstacles and len(op) > 6) else np.zeros(2)
    best = float("inf")
    for a, b, rad in _footprint_features(cfg, pose):
        if otype == 2:
            c, d = np.array([op[0], op[1]]) + t * v, np.array([op[2], op[3]]) + t * v
            dist = _seg_seg(a, b, c, d) - rad
        else:
            c = np.array([op[0], op[1]]) + t * v
            dist = _pt_seg(c, a, b) - rad - (op[4] if otype == 1 else 0.0)
        best = min(best, dist)
    return best


def obstacle_centroid(otype, op):
    return np.array([0.5 * (op[0] + op[2]), 0.5 * (op[1] + op[3])]) if otype == 2 else np.array([op[0], op[1]])


# ---- R/src/optimal_control/stage_inequality_se2.cpp:50-162 ------------------------------------------------------------------
def associate(cfg, X, obst_types, obst_params, dt):
    """Per grid point k = 1..N-2 (rows exist only for the poses _x_seq[k], k <= N-2, finite_differences_grid_se2.cpp:44-55; k = 0 is
    skipped :73) the kept obstacle indices in the reference's push_back order: forced inclusions (dist < force_inclusion_dist) in
    obstacle order, then the nearest left and the nearest right one within cutoff_dist; side = sign of cross2d(heading, centroid
    in WORLD coordinates) (:121).  Dynamic obstacles are all kept (:99-106); they come last here (own list)."""
    N = X.shape[0]
    out = [[] for _ in range(N)]
    for k in range(1, N - 1):
        pose = X[k]
        orient = np.array([math.cos(pose[2]), math.sin(pose[2])])
        left = right = -1
        lmin = rmin = float("inf")
        forced, dyn = [], []
        for j, (ot, op) in enumerate(zip(obst_types, obst_params)):
            if cfg.enable_dynamic_obstacles and (op[5] != 0.0 or op[6] != 0.0):
                dyn.append(j)
                continue
            dist = footprint_distance(cfg, pose, ot, op)
            if dist < cfg.force_inclusion_dist:
                forced.append(j)
                continue
            if dist > cfg.cutoff_dist:
                continue
            c = obstacle_centroid(ot, op)
            if orient[0] * c[1] - c[0] * orient[1] > 0:
                if dist < lmin:
                    lmin, left = dist, j
            else:
                if dist < rmin:
                    rmin, right = dist, j
        keep = list(forced)
        if left >= 0:
            keep.append(left)
        if right >= 0:
            keep.append(right)
        out[k] = keep + dyn
    return out


# ---- R/src/optimal_control/min_time_via_points_cost.cpp:40-118 + findClosestPose (full_discretization_grid_base_se2.cpp:364-388)
def associate_viapoints(cfg, X, vps):
    N = X.shape[0]
    stage = []
    start = 0
    for vp in vps:
        best, idx = float("inf"), -1
        for i in range(start, N - 1):          # _x_seq holds x_0 .. x_{N-2}
            d = math.hypot(vp[0] - X[i, 0], vp[1] - X[i, 1])
            if d < best:
                best, idx = d, i
        d = math.hypot(vp[0] - X[N - 1, 0], vp[1] - X[N - 1, 1])   # _xf
        if d < best:
            best, idx = d, N - 1
        if cfg.vp_ordered:
            start = idx + 2
        if idx > N - 2:
            idx = N - 2
        if idx < 1:
            idx = 1 if cfg.vp_ordered else -1
        stage.append(idx)
    return stage


# ---- objective: quadratic_cost_se2.cpp:31-84, final_state_conditions_se2.cpp:31-52, min_time_via_points_cost.cpp:120-145 ----
def _quad(M, d):
    M = np.array(M).reshape(len(d), len(d))
    return float(d @ M @ d)


def objective(cfg, X, U, dt, xf, vps=(), vp_stage=()):
    N = X.shape[0]
    J = 0.0

    def lx(k):
        d = X[k] - xf
        d = np.array([d[0], d[1], normalize_theta(d[2])])
        return _quad(cfg.Q, d)
    quad = cfg.objective == 1
    qz = all(q == 0.0 for q in cfg.Q); rz = all(r == 0.0 for r in cfg.R)
    hybrid = quad and cfg.hybrid_cost_minimum_time and qz and not rz    # src/controller.cpp:595-620
    if quad:
        for k in range(N - 1):
            lu = _quad(cfg.R, U[k])
            if not cfg.quadratic_integral_form:
                J += lx(k) + lu                      # finite_differences_grid_se2.cpp:47-55 (non-integral edges per stage)
            elif cfg.cost_integration == 0:
                J += dt * (lx(k) + lu)               # LeftSumCostEdge :66-70
            else:
                J += 0.5 * dt * ((lx(k) + lu) + (lx(k + 1) + lu))   # TrapezoidalIntegralCostEdge :59-65 (both ends use u_k)
    if cfg.objective in (0, 2) or hybrid:
        J += (N - 1) * dt                            # min_time_via_points_cost.cpp:52-56,120-123 / corbo::MinimumTime
    if cfg.terminal_cost and not all(cfg.xf_fixed):
        d = X[N - 1] - xf
        J += _quad(cfg.Qf, np.array([d[0], d[1], normalize_theta(d[2])]))   # final_state_conditions_se2.cpp:36-51
    if cfg.objective == 2 or (quad and cfg.vp_attraction_with_quadratic):
        for vp, k in zip(vps, vp_stage):
            if k < 0:
                continue
            J += cfg.vp_position_weight * ((vp[0] - X[k, 0]) ** 2 + (vp[1] - X[k, 1]) ** 2)
            if cfg.vp_orientation_weight > 0:
                J += cfg.vp_orientation_weight * normalize_theta(vp[2] - X[k, 2])    # linear, as coded (:142)
    return J


# ---- inequality rows c <= 0: stage_inequality_se2.cpp:164-222, final_state_conditions_se2.cpp:54-64 --------------------------
def inequality_rows(cfg, X, U, dt, xf, u_prev, u_prev_dt, assoc, obst_types, obst_params):
    N = X.shape[0]
    rows = []
    for k in range(1, N - 1):
        for j in assoc[k]:
            rows.append(cfg.min_obstacle_dist - footprint_distance(cfg, X[k], obst_types[j], obst_params[j], t=k * dt))
    lb = [cfg.du_lb[i] for i in range(2)]
    ub = [cfg.du_ub[i] for i in range(2)]
    for k in range(N):                                   # k = N-1: final control deviation edges with u_ref = 0 (:146-151)
        uk = U[k] if k <= N - 2 else np.zeros(2)
        um = U[k - 1] if k >= 1 else np.array(u_prev)
        T = dt if k >= 1 else u_prev_dt
        if k == 0 and T == 0:
            continue                                     # all-zero rows (:197-201)
        for i in range(2):
            if lb[i] > -INF:
                rows.append(lb[i] - (uk[i] - um[i]) / T)
            if ub[i] < INF:
                rows.append((uk[i] - um[i]) / T - ub[i])
    if cfg.terminal_ball and not all(cfg.xf_fixed):
        d = X[N - 1] - xf
        d = np.array([d[0], d[1], normalize_theta(d[2])])
        rows.append(_quad(cfg.terminal_ball_S, d) - cfg.terminal_ball_gamma)
    return np.array(rows)


def defects(cfg, X, U, dt):
    N = X.shape[0]
    return np.concatenate([defect(cfg, X[k], U[k], X[k + 1], dt) for k in range(N - 1)])


# ---- cold initial guess: full_discretization_grid_base_se2.cpp:192-239 with a two-pose plan (src/controller.cpp:807-857) ------
def initial_guess(cfg, x0, xf):
    N = cfg.n
    X = np.zeros((N, 3))
    for k in range(N):
        f = k / (N - 1)
        X[k] = [x0[0] + f * (xf[0] - x0[0]), x0[1] + f * (xf[1] - x0[1]), interpolate_angle(x0[2], xf[2], f)]
    X[0] = x0
    X[N - 1] = xf
    return X, np.zeros((N - 1, 2)), cfg.dt_ref


class Problem:
    """One instance: decision vector z = (unfixed states x_1..x_{N-1}, controls u_0..u_{N-2}, dt if free) in the reference's
    order (computeActiveVertices, full_discretization_grid_base_se2.cpp:564-577: per k the state then the control)."""

    def __init__(self, cfg, x0, xf, u_prev, u_prev_dt, obst_types, obst_params, vps):
        self.cfg, self.x0, self.xf = cfg, np.array(x0, float), np.array(xf, float)
        self.u_prev, self.u_prev_dt = np.array(u_prev, float), float(u_prev_dt)
        self.ot, self.op, self.vps = list(obst_types), [np.array(p, float) for p in obst_params], [np.array(v, float) for v in vps]
        self.N = cfg.n
        self.X0, self.U0, self.dt0 = initial_guess(cfg, self.x0, self.xf)
        self.assoc = associate(cfg, self.X0, self.ot, self.op, self.dt0)
        self.vp_stage = associate_viapoints(cfg, self.X0, self.vps) if self.vps else []
        self.free_xf = [i for i in range(3) if not cfg.xf_fixed[i]]

    def unpack(self, z):
        N = self.N
        X = np.zeros((N, 3)); U = np.zeros((N - 1, 2))
        X[0] = self.x0
        X[N - 1] = self.xf
        p = 0
        U[0] = z[p:p + 2]; p += 2
        for k in range(1, N - 1):
            X[k] = z[p:p + 3]; p += 3
            U[k] = z[p:p + 2]; p += 2
        for i in self.free_xf:
            X[N - 1, i] = z[p]; p += 1
        dt = z[p] if self.cfg.variable_dt else self.cfg.dt_ref
        return X, U, dt

    def pack(self, X, U, dt):
        z = list(U[0])
        for k in range(1, self.N - 1):
            z += list(X[k]) + list(U[k])
        z += [X[self.N - 1, i] for i in self.free_xf]
        if self.cfg.variable_dt:
            z.append(dt)
        return np.array(z)

    def bounds(self):
        lo, hi = [], []
        ul = [self.cfg.u_lb[i] if self.cfg.u_lb[i] > -INF else None for i in range(2)]
        uh = [self.cfg.u_ub[i] if self.cfg.u_ub[i] < INF else None for i in range(2)]
        lo += ul; hi += uh
        for k in range(1, self.N - 1):
            lo += [None] * 3 + ul; hi += [None] * 3 + uh
        lo += [None] * len(self.free_xf); hi += [None] * len(self.free_xf)
        if self.cfg.variable_dt:
            lo.append(max(self.cfg.dt_lb, 1e-3)); hi.append(self.cfg.dt_ub)   # dt > 0: the reference divides by dt (quirk 11)
        return list(zip(lo, hi))

    def f(self, z):
        X, U, dt = self.unpack(z)
        return objective(self.cfg, X, U, dt, self.xf, self.vps, self.vp_stage)

    def ceq(self, z):
        X, U, dt = self.unpack(z)
        return defects(self.cfg, X, U, dt)

    def cin(self, z):
        X, U, dt = self.unpack(z)
        return -inequality_rows(self.cfg, X, U, dt, self.xf, self.u_prev, self.u_prev_dt, self.assoc, self.ot, self.op)

    def solve_slsqp(self, maxiter=600):
        from scipy.optimize import minimize
        z0 = self.pack(self.X0, self.U0, self.dt0)
        res = minimize(self.f, z0, method="SLSQP", bounds=self.bounds(),
                       constraints=[{"type": "eq", "fun": self.ceq}, {"type": "ineq", "fun": self.cin}],
                       options=dict(maxiter=maxiter, ftol=1e-12))
        X, U, dt = self.unpack(res.x)
        cin = self.cin(res.x)
        return dict(f=float(res.fun), ceq=float(np.abs(self.ceq(res.x)).max()), cin=float(min(cin.min(), 0.0)) if len(cin) else 0.0,
                    nit=int(res.nit), U=U.tolist(), dt=float(dt), xN=X[self.N - 1].tolist())


def problem_from_batch(cfg, data, b):
    cnt = typ = par = None
    if data.get("obstacles") is not None:
        cnt, typ, par = data["obstacles"]
    ot = list(typ[b, :cnt[b]]) if cnt is not None else []
    op = [par[b, j] for j in range(cnt[b])] if cnt is not None else []
    vps = []
    if data.get("viapoints") is not None:
        vc, vposes = data["viapoints"]
        vps = [vposes[b, j] for j in range(vc[b])]
    return Problem(cfg, data["x0"][b], data["xf"][b], data["u_prev"][b], data["u_prev_dt"], ot, op, vps)
