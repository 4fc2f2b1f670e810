"""ctypes binding of the CPU ORACLE (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY: may be imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never by the product package."""
import ctypes as C
import os
import subprocess

import numpy as np

from mpc_local_planner_b200 import capi

_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_DIR, "liboracle.so")
_LIB = None


class Problem(C.Structure):
    _fields_ = [("cfg", C.POINTER(capi.Config)), ("x0", C.c_double * 3), ("xf", C.c_double * 3),
                ("u_prev", C.c_double * 2), ("u_prev_dt", C.c_double), ("n_obst", C.c_int),
                ("obst_type", C.POINTER(C.c_int)), ("obst_params", C.POINTER(C.c_double)), ("n_vp", C.c_int),
                ("vp", C.POINTER(C.c_double))]


class Ws(C.Structure):
    _fields_ = [("N", C.c_int), ("K", C.c_int), ("RS", C.c_int)] + \
        [(n, C.POINTER(C.c_double)) for n in ("X", "U", "NU", "S", "LAM", "KKT", "STEP", "OBSIDX")] + \
        [("SCAL", C.c_double * capi.SCAL_WORDS), ("vp_stage", C.c_int * 64)] + \
        [(n, C.POINTER(C.c_double)) for n in ("GL", "G", "DS", "DLAM", "XT", "UT", "P", "PI", "KG", "KT")] + \
        [("gl_dt", C.c_double), ("cold", C.c_int)]


class Result(C.Structure):
    _fields_ = [("status", C.c_int), ("iters", C.c_int), ("kkt_err", C.c_double), ("objective", C.c_double),
                ("dt", C.c_double), ("n_regularised", C.c_int), ("n_backtracks", C.c_int)]


def build(force=False):
    src = os.path.join(_DIR, "mpc_oracle.c")
    if force or not os.path.exists(LIB_PATH) or (os.path.exists(src) and os.path.getmtime(LIB_PATH) < os.path.getmtime(src)):
        subprocess.check_call(["make", "-C", _DIR, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
        L.orc_normalize_theta.restype = C.c_double
        L.orc_normalize_theta.argtypes = [C.c_double]
        L.orc_interpolate_angle.restype = C.c_double
        L.orc_interpolate_angle.argtypes = [C.c_double] * 3
        L.orc_dynamics.argtypes = [C.POINTER(capi.Config), dp, dp, dp]
        L.orc_dynamics_derivs.argtypes = [C.POINTER(capi.Config), dp, dp, dp, dp, dp, dp]
        L.orc_defect_reference.argtypes = [C.POINTER(capi.Config), dp, dp, dp, C.c_double, dp]
        L.orc_defect.argtypes = [C.POINTER(capi.Config), dp, dp, dp, C.c_double, dp]
        L.orc_footprint_distance.restype = C.c_double
        L.orc_footprint_distance.argtypes = [C.POINTER(capi.Config), dp, C.c_int, dp, dp, dp]
        L.orc_objective.restype = C.c_double
        L.orc_objective.argtypes = [C.POINTER(Problem), C.POINTER(Ws), dp, dp, C.c_double]
        L.orc_ws_alloc.restype = C.POINTER(Ws)
        L.orc_costmap_obstacles.argtypes = [C.c_int, C.c_int, C.c_double, dp, C.POINTER(C.c_ubyte), dp, C.c_double, C.c_int, dp]
        L.orc_pose_trajectory_feasible.argtypes = [C.c_int, C.c_int, C.c_double, dp, C.POINTER(C.c_ubyte), dp, C.c_int, dp, C.c_int, C.c_double, C.c_double, C.c_int]
        L.orc_resample_trajectory.argtypes = [C.c_int, dp, dp, C.c_double, C.c_int, dp, dp]
        L.orc_resample_trajectory.restype = C.c_double
        L.orc_ws_alloc.argtypes = [C.c_int, C.c_int]
        L.orc_ws_free.argtypes = [C.POINTER(Ws)]
        for f in ("orc_warm_shift", "orc_associate", "orc_init_duals", "orc_eval"):
            getattr(L, f).argtypes = [C.POINTER(Problem), C.POINTER(Ws)]
            getattr(L, f).restype = None
        L.orc_init_cold.argtypes = [C.POINTER(Problem), dp, C.POINTER(Ws)]
        L.orc_init_cold.restype = None
        L.orc_kkt_solve.argtypes = [C.POINTER(Problem), C.POINTER(Ws), C.c_double]
        L.orc_solve.argtypes = [C.POINTER(Problem), C.POINTER(Ws), C.POINTER(Result)]
        L.orc_step.argtypes = [C.POINTER(Problem), C.POINTER(Ws), dp, C.c_int, dp, dp, C.POINTER(Result)]
        L.orc_step_batch.argtypes = [C.POINTER(capi.Config), C.c_int, dp, dp, dp, C.c_double,
                                     C.POINTER(capi.Obstacles), C.POINTER(capi.ViaPoints), dp, dp, dp, dp, ip, dp, ip,
                                     C.c_int]
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


class Instance:
    """One OCP instance + its oracle workspace, with numpy views on the workspace arrays."""

    def __init__(self, cfg, x0, xf, u_prev=(0.0, 0.0), u_prev_dt=0.0, obst_types=None, obst_params=None, vp=None):
        self.L = lib()
        self.cfg = cfg.copy()
        self.N, self.K = int(cfg.n), int(cfg.k_max_obstacles_per_stage)
        self.RS = 8 + self.K
        self.p = Problem()
        self.p.cfg = C.pointer(self.cfg)
        self.p.x0[:] = list(map(float, x0))
        self.p.xf[:] = list(map(float, xf))
        self.p.u_prev[:] = list(map(float, u_prev))
        self.p.u_prev_dt = float(u_prev_dt)
        self._ot = np.ascontiguousarray(obst_types if obst_types is not None else np.zeros(0), dtype=np.int32)
        self._op = np.ascontiguousarray(obst_params if obst_params is not None else np.zeros((0, 5)), dtype=np.float64)
        self.p.n_obst = int(self._ot.shape[0])
        self.p.obst_type = self._ot.ctypes.data_as(C.POINTER(C.c_int))
        self.p.obst_params = _dp(self._op)
        self._vp = np.ascontiguousarray(vp if vp is not None else np.zeros((0, 3)), dtype=np.float64)
        self.p.n_vp = int(self._vp.shape[0])
        self.p.vp = _dp(self._vp)
        self.ws = self.L.orc_ws_alloc(self.N, self.K)

    def __del__(self):
        try:
            self.L.orc_ws_free(self.ws)
        except Exception:
            pass

    def arr(self, name):
        N, RS, K = self.N, self.RS, max(self.K, 1)
        shapes = dict(X=(3, N), U=(2, N), NU=(3, N), S=(RS, N), LAM=(RS, N), KKT=(capi.KKT_WORDS, N), STEP=(8, N),
                      OBSIDX=(K, N), GL=(5, N), G=(RS, N), DS=(RS, N), DLAM=(RS, N))
        if name == "SCAL":
            return np.ctypeslib.as_array(self.ws.contents.SCAL)
        shp = shapes[name]
        return np.ctypeslib.as_array(getattr(self.ws.contents, name), shape=shp)

    def defects(self):
        """dt-multiplied collocation defects e_k of the current iterate, [3, N-1] (orc_defect per interval)"""
        X = self.arr("X"); U = self.arr("U"); dt = float(self.arr("SCAL")[capi.SC_DT])
        e = np.zeros((self.N - 1, 3))
        for k in range(self.N - 1):
            x1 = np.ascontiguousarray(X[:, k]); x2 = np.ascontiguousarray(X[:, k + 1]); u1 = np.ascontiguousarray(U[:, k])
            self.L.orc_defect(C.byref(self.cfg), _dp(x1), _dp(u1), _dp(x2), dt, _dp(e[k]))
        return e.T.copy()

    def vp_stage(self):
        return np.ctypeslib.as_array(self.ws.contents.vp_stage)[: self.p.n_vp].copy()

    def init_cold(self, x_init=None):
        xi = np.ascontiguousarray(x_init, dtype=np.float64) if x_init is not None else None
        self.L.orc_init_cold(C.byref(self.p), _dp(xi), self.ws)

    def warm_shift(self):
        self.L.orc_warm_shift(C.byref(self.p), self.ws)

    def associate(self):
        self.L.orc_associate(C.byref(self.p), self.ws)

    def init_duals(self):
        self.L.orc_init_duals(C.byref(self.p), self.ws)

    def eval(self):
        self.L.orc_eval(C.byref(self.p), self.ws)

    def kkt_solve(self, delta=0.0):
        return self.L.orc_kkt_solve(C.byref(self.p), self.ws, float(delta))

    def solve(self):
        r = Result()
        self.L.orc_solve(C.byref(self.p), self.ws, C.byref(r))
        return r

    def step(self, x_init=None, reinit=False):
        N = self.N
        u = np.empty((N, 2))
        x = np.empty((N, 3))
        r = Result()
        xi = np.ascontiguousarray(x_init, dtype=np.float64) if x_init is not None else None
        self.L.orc_step(C.byref(self.p), self.ws, _dp(xi), 1 if reinit else 0, _dp(u), _dp(x), C.byref(r))
        return u, x, r

    def set_measurement(self, x0, xf=None, u_prev=None, u_prev_dt=None):
        self.p.x0[:] = list(map(float, x0))
        if xf is not None:
            self.p.xf[:] = list(map(float, xf))
        if u_prev is not None:
            self.p.u_prev[:] = list(map(float, u_prev))
        if u_prev_dt is not None:
            self.p.u_prev_dt = float(u_prev_dt)

    def objective(self, X, U, dt):
        X = np.ascontiguousarray(X)
        U = np.ascontiguousarray(U)
        return self.L.orc_objective(C.byref(self.p), self.ws, _dp(X), _dp(U), float(dt))


def costmap_obstacles(cost, origin, resolution, robot_pose, behind_dist, max_out):
    """updateObstacleContainerWithCostmap for one robot: cost [size_y, size_x] uint8 -> (xy [min(found, max_out), 2], found)"""
    cost = np.ascontiguousarray(cost, dtype=np.uint8)
    origin = np.ascontiguousarray(origin, dtype=np.float64); pose = np.ascontiguousarray(robot_pose, dtype=np.float64)
    xy = np.zeros((max_out, 2))
    found = lib().orc_costmap_obstacles(cost.shape[1], cost.shape[0], float(resolution), _dp(origin),
                                        cost.ctypes.data_as(C.POINTER(C.c_ubyte)), _dp(pose), float(behind_dist), int(max_out), _dp(xy))
    return xy[: min(found, max_out)].copy(), found


def pose_trajectory_feasible(cost, origin, resolution, x_seq, footprint, inscribed_radius, min_resolution_angular, look_ahead_idx=-1):
    """isPoseTrajectoryFeasible for one robot: cost [size_y, size_x] uint8, x_seq [n, 3], footprint [n_fp, 2] -> bool"""
    cost = np.ascontiguousarray(cost, dtype=np.uint8); origin = np.ascontiguousarray(origin, dtype=np.float64)
    x_seq = np.ascontiguousarray(x_seq, dtype=np.float64); fp = np.ascontiguousarray(footprint, dtype=np.float64).reshape(-1, 2)
    return bool(lib().orc_pose_trajectory_feasible(cost.shape[1], cost.shape[0], float(resolution), _dp(origin), cost.ctypes.data_as(C.POINTER(C.c_ubyte)),
                                                   _dp(x_seq), x_seq.shape[0], _dp(fp), fp.shape[0], float(inscribed_radius),
                                                   float(min_resolution_angular), int(look_ahead_idx)))


def resample_trajectory(X, U, dt, n_new):
    """resampleTrajectory(n_new) of one trajectory: X [3, n], U [2, n] (component-major) -> (Xn [3, n_new], Un [2, n_new], dt_new)"""
    X = np.ascontiguousarray(X, dtype=np.float64); U = np.ascontiguousarray(U, dtype=np.float64)
    n = X.shape[1]
    Xn = np.zeros((3, n_new)); Un = np.zeros((2, n_new))
    dt_new = lib().orc_resample_trajectory(n, _dp(X), _dp(U), float(dt), int(n_new), _dp(Xn), _dp(Un))
    return Xn, Un, dt_new


def instance_from_batch(cfg, data, b):
    cnt, types, params = data["obstacles"] if data.get("obstacles") is not None else (None, None, None)
    ot = types[b, : cnt[b]] if cnt is not None else None
    op = params[b, : cnt[b]] if cnt is not None else None
    vp = None
    if data.get("viapoints") is not None:
        vc, vposes = data["viapoints"]
        vp = vposes[b, : vc[b]]
    return Instance(cfg, data["x0"][b], data["xf"][b], data["u_prev"][b], data["u_prev_dt"], ot, op, vp)


def step_batch(cfg, data, n_threads=1, x_init=None):
    """Cold Controller::step over a batch on the CPU oracle; same outputs as capi.BatchSolver.step."""
    L = lib()
    x0 = np.ascontiguousarray(data["x0"], dtype=np.float64)
    xf = np.ascontiguousarray(data["xf"], dtype=np.float64)
    up = np.ascontiguousarray(data["u_prev"], dtype=np.float64)
    B, N = x0.shape[0], int(cfg.n)
    o = v = None
    keep = []
    if data.get("obstacles") is not None:
        o, k = capi.pack_obstacles(*data["obstacles"])
        keep.append(k)
    if data.get("viapoints") is not None:
        v, k = capi.pack_viapoints(*data["viapoints"])
        keep.append(k)
    out = dict(u_seq=np.empty((B, N, 2)), x_seq=np.empty((B, N, 3)), dt=np.empty(B),
               status=np.empty(B, dtype=np.int32), kkt_err=np.empty(B), iters=np.empty(B, dtype=np.int32))
    xi = np.ascontiguousarray(x_init, dtype=np.float64) if x_init is not None else None
    c = cfg.copy()
    L.orc_step_batch(C.byref(c), B, _dp(x0), _dp(xf), _dp(up), float(data["u_prev_dt"]),
                     C.byref(o) if o else None, C.byref(v) if v else None, _dp(xi), _dp(out["u_seq"]),
                     _dp(out["x_seq"]), _dp(out["dt"]), out["status"].ctypes.data_as(C.POINTER(C.c_int)),
                     _dp(out["kkt_err"]), out["iters"].ctypes.data_as(C.POINTER(C.c_int)), int(n_threads))
    return out
