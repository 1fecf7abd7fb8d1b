#!/usr/bin/env python
"""Proof sweep (run on the GPU box): candidates of the committed cfg4 forest corridors (whole sweeps, N=10, 3 polytopes from
JPS3D + decomposition; safe sweeps, 4 polytopes, free final position) and of synthetic cfg5 corridors (N=15, 8 polytopes) go
through the product kernel (fq_solve_batch, coefficients) and the certifying kernel (fq_solve_batch_cert), and EVERY sampled
flag is then proved on the literal rows of the reference's model (oracle/model_fullspace.py) with oracle/proofs.py:
  solved      -> primal feasibility of the GPU's coefficients + KKT multipliers (optimality), reported cost = cost of the point;
  not solved  -> the exported Farkas certificate.
No CPU solver's verdict enters: a failure would be an assertion, not a mismatch count.  Prints one JSON line.
usage: stress_proofs.py [corridors=48] [candidates per corridor and sweep=48]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                                 # noqa: E402
from faster_b200 import capi, corridor as cr                 # noqa: E402
from oracle import model_fullspace as mf, proofs             # noqa: E402

n_corr = int(sys.argv[1]) if len(sys.argv) > 1 else 48
n_cand = int(sys.argv[2]) if len(sys.argv) > 2 else 48
s = capi.Solver(0)
rng = np.random.default_rng(20260923)
res = {"corridors": 0, "candidates": 0, "solved_proved_optimal": 0, "not_solved_proved_infeasible": 0, "kernels_disagree": 0,
       "abandoned": 0, "worst_stationarity_resid": 0.0, "worst_gap": 0.0, "worst_row_excess": 0.0, "with_tight_rows": 0,
       "proof_failures": 0, "first_failures": [], "by_kind": {}}
t0 = time.time()


def polys_of(w, kind, j):
    po, fo, Ab = w["poly_ofs_" + kind], w["face_ofs_" + kind], w["Ab_" + kind]
    return [(Ab[fo[p]:fo[p + 1], :3].copy(), Ab[fo[p]:fo[p + 1], 3].copy()) for p in range(po[j], po[j + 1])]


def prove_batch(kind, N, x0, xf, lim, polys, dts, sigs, ff):
    fg, cg, cog, _ = s.solve_batch(N, x0, xf, lim, polys, dts, sigs, ff, want_coeffs=True)
    fc, _, cert = s.solve_batch_cert(N, x0, xf, lim, polys, dts, sigs, ff)
    res["kernels_disagree"] += int((fg != fc).sum())
    k = res["by_kind"].setdefault(kind, {"solved": 0, "not_solved": 0})
    for i in range(len(dts)):
        model = mf.build(N, x0, xf, lim, dts[i], polys, sigs[i], ff)
        res["candidates"] += 1
        try:
            prove_one(kind, k, N, polys, sigs[i], model, fg[i], cg[i], cog[i], cert[i])
        except AssertionError as e:                         # counted and shown, never swallowed: the run fails at the end
            res["proof_failures"] += 1
            if len(res["first_failures"]) < 5:
                res["first_failures"].append("%s dt=%.6g sigma=%s flag=%d: %s" % (kind, dts[i], list(map(int, sigs[i])), fg[i], str(e)[:300]))


def prove_one(kind, k, N, polys, sigma, model, flag, cost, coeffs, cert_row):
    if flag:
        r = proofs.assert_optimal(model, coeffs, cost)
        res["solved_proved_optimal"] += 1
        k["solved"] += 1
        res["with_tight_rows"] += int(r["n_active"] > 0)
        res["worst_stationarity_resid"] = max(res["worst_stationarity_resid"], r["resid"])
        res["worst_gap"] = max(res["worst_gap"], r["gap"])
        res["worst_row_excess"] = max(res["worst_row_excess"], r["eq"], r["ineq"])
    elif int(cert_row[0]) >= 1:
        proofs.assert_infeasible(model, N, polys, sigma, cert_row)
        res["not_solved_proved_infeasible"] += 1
        k["not_solved"] += 1
    else:
        res["abandoned"] += 1                                # iteration cap / non-finite input: no verdict, no proof


w = bench.load_cfg4(0, n_corr)
g = s.replan_pairs(w)                                        # the chain supplies each corridor's dt bases and R (safe x0)
rr = g["results"]
for j in range(n_corr):
    for kind, N, ff in (("whole", w["N_whole"], True), ("safe", w["N_safe"], False)):
        if kind == "safe" and not np.isfinite(rr["safe_dt_base"][j]):
            continue                                         # no whole trajectory: the reference returns before the safe sweep
        fac, sig_all = w["factors_" + kind], w["sigmas_" + kind]
        base = rr[kind + "_dt_base"][j]
        pick_f = rng.integers(0, len(fac), n_cand)
        pick_s = rng.integers(0, len(sig_all), n_cand)
        x0 = w["x0"][j] if kind == "whole" else rr["R"][j]
        prove_batch("cfg4_" + kind, N, x0, w["xf_" + kind][j], w["lim"][j], polys_of(w, kind, j), fac[pick_f] * base, sig_all[pick_s], ff)
    res["corridors"] += 1
for seed in range(5000, 5000 + max(2, n_corr // 8)):        # BASELINE config 5: ground robot, N=15, 8 narrow polytopes
    pb = cr.make_corridor(seed, 8, 15, "ground", True)
    allm = cr.sample_monotone_sigmas(15, 8, 256, rng)
    base = max(capi.dt_initial(pb["x0"], pb["xf"], pb["lim"], 15), 0.02)
    prove_batch("cfg5", 15, pb["x0"], pb["xf"], pb["lim"], pb["polys"], rng.choice(np.arange(1, 17), n_cand) * base,
                allm[rng.integers(0, len(allm), n_cand)], True)
    res["corridors"] += 1
res["seconds"] = round(time.time() - t0, 1)
print(json.dumps(res))
assert res["kernels_disagree"] == 0 and res["proof_failures"] == 0
