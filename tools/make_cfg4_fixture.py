#!/usr/bin/env python
"""Generates bench_data/cfg4_forest.npz: the corridors of BASELINE config 4 (random forest, JPS3D + convex decomposition
on the host, paired whole + safe problems).  Run here (CPU only); both arms of bench.py load the file, so they see the
same inputs and neither needs the other's library to make them.

Per corridor: whole problem from faster_b200.corridor.make_forest_pair_whole (the product's host-side JPS3D and
decomposition); R = sample (int)(0.6 n) of the whole winner, computed with the CPU restatement (oracle) over the very
candidate grid the bench uses; safe problem from make_forest_pair_safe(R).  The GPU chain recomputes R itself; the R
stored here only shaped the safe corridor.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from faster_b200 import capi, corridor as cr        # noqa: E402
from oracle import pyoracle as po                   # noqa: E402

N, PW, PS, N_FAC, N_SIG, DC, RFRAC = 10, 3, 4, 16, 64, 0.01, 0.6


def grid_sigmas(P):
    allm = cr.monotone_sigmas(N, P)
    idx = np.linspace(0, len(allm) - 1, N_SIG).round().astype(int)
    return allm[idx]


def main(n_corr=512, seed0=40000):
    sw, ss = grid_sigmas(PW), grid_sigmas(PS)
    whole, safe, Rs = [], [], []
    seed = seed0
    while len(whole) < n_corr:
        seed += 1
        try:
            w = cr.make_forest_pair_whole(seed, N, PW)
        except RuntimeError:
            continue
        g = po.gen_new_traj(N, w["x0"], w["xf"], w["lim"], w["polys"], DC, 1.0, float(N_FAC), 1.0, sw, True)
        if not g["solved"]:
            continue                                   # a replan without whole trajectory has no safe problem
        X = po.fill_x(N, g["coeffs"], g["dt"], DC)
        k = min(len(X) - 1, int(RFRAC * len(X)))
        R = X[k, :9].copy()
        s = cr.make_forest_pair_safe(w, R, N, PS)
        whole.append(w); safe.append(s); Rs.append(R)
    pk = capi.make_pair_workload(whole, safe, np.arange(1.0, N_FAC + 1), sw, np.arange(1.0, N_FAC + 1), ss, DC, RFRAC)
    out = os.path.join(ROOT, "bench_data", "cfg4_forest.npz")
    np.savez_compressed(out, R_oracle=np.array(Rs), seeds=np.array([w["seed"] for w in whole]),
                        **{k: v for k, v in pk.items() if isinstance(v, np.ndarray)},
                        meta=np.array([pk["n_prob"], pk["N_whole"], pk["N_safe"], pk["max_faces_whole"], pk["max_poly_faces_whole"],
                                       pk["max_faces_safe"], pk["max_poly_faces_safe"]]), DC=DC, r_fraction=RFRAC)
    print(out, os.path.getsize(out), "bytes;", n_corr, "corridors; faces/polytope whole %.1f safe %.1f" %
          (np.diff(pk["face_ofs_whole"]).mean(), np.diff(pk["face_ofs_safe"]).mean()))


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 512)
