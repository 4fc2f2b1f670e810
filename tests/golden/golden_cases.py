"""The cases of the numpy/SLSQP golden fixtures: configuration + seeded instances per case name.  Shared by the generator
(tests/golden/make_golden.py) and by the tests that check the oracle (CPU) and the CUDA path (GPU) against the fixtures."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from mpc_local_planner_b200 import capi, configs  # noqa: E402

CASES = ("g1", "cfg2", "cfg4", "cfg2_midpoint", "cfg2_trapezoidal", "cfg2_circular_footprint", "cfg1_obstacles")


def make(name, tol=1e-9):
    """-> (config, data, rows wanted, instances to try).  SLSQP always runs from the reference's cold initial guess."""
    n = 24
    if name == "g1":          # the reference's only fixed scenario (src/test_mpc_optim_node.cpp:67-69,105-106), N = 20
        cfg, data, want, pool = configs.cfg1(tol=tol), configs.g1_instance(), 1, 1
    elif name == "cfg2":
        cfg, data, want, pool = configs.cfg2(n=n, tol=tol), configs.generate(5, 40, n=n), 10, 40
    elif name == "cfg4":      # via-point attraction added to the quadratic form (SURVEY 8d cfg 4, reading A)
        cfg = configs.cfg4(n=n, tol=tol)
        data = configs.generate(5, 40, n=n)
        rng = np.random.default_rng(4)
        vp = np.zeros((40, 2, 3))
        for b in range(40):
            d = data["xf"][b, :2]
            nrm = np.array([-d[1], d[0]]) / np.linalg.norm(d)
            for j, fr in enumerate((1 / 3, 2 / 3)):
                vp[b, j, :2] = fr * d + rng.uniform(-0.3, 0.3) * nrm
        data["viapoints"] = (np.full(40, 2, dtype=np.int32), vp)
        want, pool = 6, 40
    elif name == "cfg2_midpoint":
        cfg = configs.cfg2(n=n, tol=tol); cfg.collocation = capi.COLLOC_MIDPOINT
        data, want, pool = configs.generate(5, 32, n=n), 5, 32
    elif name == "cfg2_trapezoidal":
        cfg = configs.cfg2_trapezoidal(n=n, tol=tol, variable_dt=False)
        data, want, pool = configs.generate(5, 32, n=n), 4, 32
    elif name == "cfg2_circular_footprint":
        cfg = configs.cfg2(n=n, tol=tol)
        cfg.footprint_type = capi.FOOTPRINT_CIRCULAR; cfg.footprint_params[0] = 0.12; cfg.min_obstacle_dist = 0.1
        data, want, pool = configs.generate(5, 32, n=n), 4, 32
    elif name == "cfg1_obstacles":   # minimum time, fixed goal, free dt, obstacles that matter (cfg-2 style instances at N = 20)
        cfg = configs.cfg1(tol=tol)
        cfg.min_obstacle_dist = 0.2; cfg.k_max_obstacles_per_stage = 5
        cfg.initial_guess_bumps = 0   # the solver's start stays on the straight line SLSQP starts from (same homotopy class)
        data, want, pool = configs.generate(5, 24, n=20), 5, 24
    else:
        raise KeyError(name)
    # The solver under test starts from the reference's cold guess as well, except on the free-dt / fixed-goal cases: with zero
    # controls the linearised dynamics of a nonholonomic robot cannot move sideways, the linearised fixed-goal constraints are
    # rank deficient, and the interior-point method here has no constraint regularisation (Ipopt's delta_c) -- those cases run from
    # the solver's seeded start (the minimum time does not depend on the start unless obstacles split the homotopy classes).
    cfg.reference_initial_guess = 0 if cfg.variable_dt else 1
    return cfg, data, want, pool
