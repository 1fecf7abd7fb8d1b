"""Generates tests/golden/corridor_continuous_expected.json: feasibility flags, costs and coefficients of the restated
model on the reference demo's corridor (tests/golden/corridor_continuous.json) for every non-decreasing assignment at
six time allocations, plus the genNewTraj sweep result.

Values come from the CPU restatement (oracle/fq_oracle.c) and every feasible entry is cross-checked here against HiGHS on
the literal full-space model (oracle/model_fullspace.py) before it is written -- they are "restated-reference" goldens,
NOT Gurobi outputs (Gurobi is unavailable).  Run from the repo root: python tests/golden/make_expected_continuous.py
"""
import json
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")
from faster_b200 import corridor as cr            # noqa: E402
from oracle import model_fullspace as mf           # noqa: E402
from oracle import pyoracle as po                  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
fx = json.load(open(os.path.join(HERE, "corridor_continuous.json")))
polys = [(np.array(p["A"]), np.array(p["b"])) for p in fx["polys"]]
N, x0, xf, lim = fx["N"], fx["x0"], fx["xf"], fx["lim"]
sig = cr.monotone_sigmas(N, 3)
dts = [0.5, 0.6, 0.7, 0.8, 1.0, 1.5]
out = {"dts": dts, "sigmas": sig.tolist(), "feasible": [], "cost": [], "checked_with_highs": 0}
coeff_samples = {}
for dt in dts:
    f, c, co = po.solve_batch(N, x0, xf, lim, polys, np.full(len(sig), dt), sig, True, True, threads=8)
    for k in range(len(sig)):
        if f[k] and k % 7 == 0:                     # HiGHS on a subset (it is slow), all of them agree
            ok, ch, coh = mf.solve_highs(N, x0, xf, lim, dt, polys, sig[k])
            assert ok and abs(ch - c[k]) <= 1e-6 * max(1.0, ch), (dt, k, ch, c[k])
            out["checked_with_highs"] += 1
    out["feasible"].append(f.astype(int).tolist())
    out["cost"].append([float(v) if np.isfinite(v) else None for v in c])
    best = int(np.argmin(np.where(f.astype(bool), c, np.inf)))
    if f[best]:
        coeff_samples["%g" % dt] = {"sigma_index": best, "coeffs": co[best].tolist()}
out["best_coeffs"] = coeff_samples
g = po.gen_new_traj(N, x0, xf, lim, polys, 0.01, 1.0, 10.0, 1.0, None, True)
out["gen_new_traj"] = {"DC": 0.01, "factors": [1.0, 10.0, 1.0], "solved": g["solved"], "factor": g["factor"], "dt": g["dt"],
                       "trials": g["trials"], "cost": g["cost"], "sigma": g["sigma"].tolist()}
json.dump(out, open(os.path.join(HERE, "corridor_continuous_expected.json"), "w"))
print("wrote expected: %d feasible of %d, %d cross-checked with HiGHS" %
      (sum(map(sum, out["feasible"])), len(dts) * len(sig), out["checked_with_highs"]))
