"""The C-ABI shared library on a box WITHOUT a GPU: it loads, exports every symbol of include/mpcb200.h, agrees with the
Python twin of the config struct, and refuses to run (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from mpc_local_planner_b200 import capi
from tests.conftest import ROOT, has_gpu


def test_header_symbols_exported(cuda_lib):
    hdr = open(os.path.join(ROOT, "include", "mpcb200.h")).read()
    declared = set(re.findall(r"\b(mpcb200_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"mpcb200_handle"}
    assert declared, "no declarations found"
    for sym in declared:
        assert hasattr(cuda_lib, sym), f"{sym} declared in include/mpcb200.h but not exported"
    assert declared == set(capi.EXPORTS), declared ^ set(capi.EXPORTS)


def test_default_config_matches_python_twin(cuda_lib):
    c = capi.Config()
    cuda_lib.mpcb200_default_config(C.byref(c))
    p = capi.default_config()
    assert bytes(c) == bytes(p), "struct layout or defaults differ between mpcb200.h and capi.Config"
    # in-code defaults of the reference
    assert (c.n, c.dt_ref, c.max_iter, c.objective) == (20, 0.3, 100, capi.OBJ_MINIMUM_TIME)
    assert list(c.u_lb) == [-0.2, -0.3] and list(c.u_ub) == [0.4, 0.3]
    assert (c.min_obstacle_dist, c.force_inclusion_dist, c.cutoff_dist) == (0.5, 0.5, 2.0)


def test_invalid_configs_rejected(cuda_lib):
    h = C.c_void_p()
    c = capi.default_config(); c.n = 2
    assert cuda_lib.mpcb200_create(C.byref(c), 4, 0, C.byref(h)) == capi.E_INVALID
    c = capi.default_config(); c.collocation = capi.COLLOC_CRANK_NICOLSON
    assert cuda_lib.mpcb200_create(C.byref(c), 4, 0, C.byref(h)) == capi.E_UNSUPPORTED
    assert b"forward_differences" in cuda_lib.mpcb200_last_error(None)
    c = capi.default_config(); c.objective = capi.OBJ_MINIMUM_TIME; c.variable_dt = 0
    assert cuda_lib.mpcb200_create(C.byref(c), 4, 0, C.byref(h)) == capi.E_INVALID


@pytest.mark.skipif(has_gpu(), reason="box has a GPU")
def test_no_cpu_fallback(cuda_lib):
    """Without a CUDA device the product path fails loudly."""
    with pytest.raises(capi.SolverError) as ei:
        capi.BatchSolver(capi.default_config(), 4)
    assert "no CPU fallback" in str(ei.value) or "CUDA" in str(ei.value)


@pytest.mark.skipif(has_gpu(), reason="box has a GPU")
def test_multi_device_handle_fails_loudly_without_gpu(cuda_lib):
    from mpc_local_planner_b200 import capi, configs
    with pytest.raises(capi.SolverError) as e:
        capi.MultiSolver(configs.cfg2(), 64, [0, 1])
    assert "no CPU fallback" in str(e.value)
