"""mpc_local_planner_b200 -- B200-native batched receding-horizon OCP solver behind the Controller::step() surface of
rst-tu-dortmund/mpc_local_planner.  The compute path is hand-written sm_100a CUDA in ``csrc/`` behind the C ABI of
``include/mpcb200.h``; this package only holds the ctypes binding (``capi``) and the benchmark/test workload
definitions (``configs``).  No CPU fallback exists in the product path."""
from . import capi, configs  # noqa: F401

__all__ = ["capi", "configs"]
