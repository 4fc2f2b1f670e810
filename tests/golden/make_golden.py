#!/usr/bin/env python
"""Generates the golden fixtures of tests/golden/ (committed; this script is the provenance).

The reference ships no golden vectors and cannot be run here.  The known answers are therefore produced by an
INDEPENDENT ALGORITHM ON AN INDEPENDENT RESTATEMENT: scipy.optimize SLSQP with numeric Jacobians on tests/golden/ocp_numpy.py --
objective, collocation defects, obstacle association, obstacle / control-rate rows and the cold initial guess written in numpy
directly from the reference's source lines.  Nothing of oracle/ (nor of the CUDA library) is imported, loaded or called here.
The oracle's interior-point solver and the CUDA solver are then both tested against these fixtures (tests/test_oracle_golden.py,
tests/test_gpu_parity.py), in the "reference initial guess" mode (config.reference_initial_guess = 1: straight line, zero
controls -- the start the numpy restatement builds from full_discretization_grid_base_se2.cpp:192-239).

Horizons are short (N = 20..24) because every function evaluation is a Python loop.

    python tests/golden/make_golden.py [case ...]      # rewrites tests/golden/np_<case>.json (minutes per case)
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import golden_cases  # noqa: E402
import ocp_numpy as on  # noqa: E402


def main():
    want = sys.argv[1:] or list(golden_cases.CASES)
    for name in want:
        cfg, data, rows_wanted, pool = golden_cases.make(name)
        rows = []
        for b in range(pool):
            p = on.problem_from_batch(cfg, data, b)
            t = time.time()
            r = p.solve_slsqp()
            ok = r["ceq"] < 1e-8 and r["cin"] > -1e-8 and r["nit"] < 600
            print("%s inst %d f %.6f ceq %.1e cin %.1e nit %d dt %.6f %s (%.0fs)" % (name, b, r["f"], r["ceq"], r["cin"], r["nit"], r["dt"],
                                                                              "kept" if ok else "not converged", time.time() - t), flush=True)
            if ok:
                r["instance"] = b
                rows.append(r)
            if len(rows) >= rows_wanted:
                break
        json.dump(dict(case=name, n=int(cfg.n), pool=pool, rows=rows), open(os.path.join(HERE, "np_%s.json" % name), "w"), indent=1)


if __name__ == "__main__":
    main()
