#!/usr/bin/env python
"""Stress run of the exact MIQP path (GPU branch-and-bound) against the oracle's branch-and-bound (run on the GPU box).
usage: stress_exact.py [n_seeds]  -> one JSON line."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from faster_b200 import capi, corridor as cr          # noqa: E402
from oracle import pyoracle as po                      # noqa: E402

n_seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 60
s = capi.Solver(0)
res = {"sweeps": 0, "sweep_mismatches": 0, "single_dt_cases": 0, "single_dt_mismatches": 0, "non_monotone_optima": 0,
       "inexact_flags": 0, "max_nodes": 0, "mean_nodes": 0.0, "examples": []}
nodes = []
t0 = time.time()
for seed in range(n_seeds):
    for kind in ("synthetic", "forest"):
        for (N, P, ff) in ((10, 3, True), (10, 4, False), (6, 3, True), (8, 3, False)):
            try:
                pb = cr.make_corridor(21000 + seed, P, N, "uav", ff) if kind == "synthetic" else cr.make_forest_corridor(22000 + seed, P, N, ff)
            except RuntimeError:
                continue
            dti = capi.dt_initial(pb["x0"], pb["xf"], pb["lim"], N)
            dts = np.arange(1.0, 11.0) * max(dti, 0.02)
            g = s.gen_new_traj_exact(N, pb["x0"], pb["xf"], pb["lim"], pb["polys"], dts, ff)
            o = po.gen_new_traj(N, pb["x0"], pb["xf"], pb["lim"], pb["polys"], 0.01, 1.0, 10.0, 1.0, None, ff)
            res["sweeps"] += 1
            res["inexact_flags"] += int(not g["exact"])
            nodes.append(g["nodes"])
            bad = g["solved"] != o["solved"] or (o["solved"] and (g["dt_index"] + 1 != o["trials"] or abs(g["cost"] - o["cost"]) > 1e-7 * max(1.0, o["cost"])))
            if bad and g["exact"]:
                res["sweep_mismatches"] += 1
                res["examples"].append(["sweep", kind, seed, N, P, ff, g["solved"], g["dt_index"], g["cost"], o["solved"], o["trials"], o["cost"]])
            # one loose time allocation: where non-monotone optima live
            dt = 5.0 * max(dti, 0.02)
            g1 = s.gen_new_traj_exact(N, pb["x0"], pb["xf"], pb["lim"], pb["polys"], [dt], ff)
            rc, c, _, sg, _ = po.solve_miqp(N, pb["x0"], pb["xf"], pb["lim"], dt, pb["polys"], ff)
            res["single_dt_cases"] += 1
            res["inexact_flags"] += int(not g1["exact"])
            nodes.append(g1["nodes"])
            if g1["exact"] and (g1["solved"] != (rc == 1) or (rc == 1 and abs(g1["cost"] - c) > 1e-7 * max(1.0, c))):
                res["single_dt_mismatches"] += 1
                res["examples"].append(["single", kind, seed, N, P, ff, g1["solved"], g1["cost"], rc, c])
            if rc == 1 and np.any(np.diff(g1["sigma"].astype(int)) < 0):
                res["non_monotone_optima"] += 1
res["max_nodes"] = int(max(nodes))
res["mean_nodes"] = float(np.mean(nodes))
res["seconds"] = time.time() - t0
res["examples"] = res["examples"][:8]
print(json.dumps(res))
