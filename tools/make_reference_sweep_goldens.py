#!/usr/bin/env python
"""Golden vectors from THE REFERENCE ITSELF (run where /root/reference exists): the reference's own genNewTraj
(faster/src/solverGurobi.cpp compiled unmodified into oracle/_ref/libsolver_ref.so over the recording Gurobi stand-in), with
HiGHS + enumeration of the binaries answering optimize(), on seeded synthetic corridors.  Writes
tests/golden/reference_sweeps.json: inputs (x0, xf, limits, polytopes, factor window) and what the reference's loop returned
(solved, trials_, dt_, factor_that_worked_, coefficients, the first and last sampled states of fillX).  The fixture travels to
boxes without /root/reference; tests/test_reference_solver_cpu.py checks the CPU restatement against it there."""
import json
import os
import sys

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from faster_b200 import corridor as cr                         # noqa: E402
from oracle import model_fullspace as mf, solver_ref as sr      # noqa: E402


def highs(q, Aeq, beq, Ain, bin_):
    ok, z = mf.solve_qp_highs(sp.diags(2.0 * q).tocsc(), Aeq, beq, Ain, bin_)
    return ok, z, (float(np.sum(q * z * z)) if ok else np.inf)


cases = []
for k, (N, P, ff, window, shift) in enumerate([(4, 2, True, (1.0, 6.0, 1.0), 0.0), (5, 2, True, (1.0, 6.0, 1.0), 0.0), (5, 2, False, (1.0, 5.0, 1.0), 0.0),
                                               (4, 3, True, (1.0, 6.0, 1.0), 0.0), (6, 2, True, (1.0, 4.0, 0.5), 0.0), (4, 0, True, (1.0, 6.0, 1.0), 0.0),
                                               (3, 0, True, (1.0, 10.0, 1.0), 0.0), (4, 2, True, (1.0, 3.0, 1.0), 40.0), (5, 3, False, (2.0, 6.0, 2.0), 0.0),
                                               (4, 2, False, (1.0, 6.0, 1.0), 0.0)]):
    pb = cr.make_corridor(8800 + k, max(P, 1), N, "uav", ff)
    polys = pb["polys"] if P else []
    xf = np.array(pb["xf"], float)
    xf[:3] += shift
    r = sr.gen_new_traj(N, pb["x0"], xf, pb["lim"], polys, 0.01, *window, highs, ff)
    cases.append(dict(N=N, P=P, force_final=ff, DC=0.01, window=list(window), x0=list(map(float, pb["x0"])), xf=list(map(float, xf)),
                      lim=list(map(float, pb["lim"])), polys=[dict(A=np.asarray(A).tolist(), b=np.asarray(b).tolist()) for A, b in polys],
                      solved=r["solved"], trials=r["trials"], dt=r["dt"], factor=r["factor"], n_optimize=r["n_optimize"],
                      coeffs=r["coeffs"].tolist() if r["solved"] else None, n_samples=len(r["samples"]),
                      first_sample=r["samples"][0].tolist() if r["solved"] else None,
                      last_sample=r["samples"][-1].tolist() if r["solved"] else None))
    print(k, N, P, ff, "->", r["solved"], r["trials"], r["dt"])
out = os.path.join(ROOT, "tests", "golden", "reference_sweeps.json")
json.dump(dict(generator="tools/make_reference_sweep_goldens.py", solver_in_place_of_gurobi="HiGHS (scipy-vendored) per assignment, all P^N enumerated",
               cases=cases), open(out, "w"), indent=0)
print("wrote", out, os.path.getsize(out), "bytes")
