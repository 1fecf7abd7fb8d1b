#!/usr/bin/env python
"""Large parity sweep (run on the GPU box): CUDA path vs the CPU restatement on many random corridors.
usage: stress_parity.py [n_corridors]   -> prints one JSON line (flag mismatches, worst cost error, iteration stats)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                    # noqa: E402
from faster_b200 import capi                    # noqa: E402
from oracle import pyoracle as po               # noqa: E402

n_corr = int(sys.argv[1]) if len(sys.argv) > 1 else 128
s = capi.Solver(0)
res = {"corridors_per_kind": n_corr, "candidates": 0, "flag_mismatches": 0, "max_rel_cost_err": 0.0, "feasible": 0,
       "iteration_cap_hits": 0, "mismatch_examples": []}
t0 = time.time()
for kind, seed0 in (("whole", 700000), ("safe", 800000)):
    w = bench.make_workload(n_corr, seed0, kind)
    fg, cg, _, it = s.solve_multi(w["N"], w["ff"], w["x0"], w["xf"], w["lim"], w["poly_ofs"], w["face_ofs"], w["Ab"],
                                  w["cand_ofs"], w["dt"], w["sigma"], want_iters=True)
    fo, co = po.solve_multi(w["N"], w["ff"], w["x0"], w["xf"], w["lim"], w["poly_ofs"], w["face_ofs"], w["Ab"],
                            w["cand_ofs"], w["dt"], w["sigma"], os.cpu_count() or 1)
    bad = np.nonzero(fg != fo)[0]
    res["candidates"] += int(fg.size)
    res["flag_mismatches"] += int(bad.size)
    res["feasible"] += int(fo.sum())
    res["iteration_cap_hits"] += int((it < 0).sum())
    ok = fo.astype(bool) & fg.astype(bool)
    if ok.any():
        res["max_rel_cost_err"] = max(res["max_rel_cost_err"], float((np.abs(cg[ok] - co[ok]) / np.maximum(1e-9, np.abs(co[ok]))).max()))
    for i in bad[:5]:
        res["mismatch_examples"].append({"kind": kind, "cand": int(i), "gpu": int(fg[i]), "oracle": int(fo[i]), "dt": float(w["dt"][i]),
                                         "gpu_cost": float(cg[i]), "oracle_cost": float(co[i])})
res["seconds"] = time.time() - t0
print(json.dumps(res))
