#!/usr/bin/env python
"""CPU model (numpy mirror of the kernel, tests/kernel_mirror.py): how much of the corridor-row scan could an EXACT distance
bound skip?  The rank of a row is its signed distance to the iterate in w-space (unit face normals), so between two scans it
moves by at most the length of the path w travelled.  Policy modelled: the first scan of a candidate is complete and splits
the items into a NEAR list (the `near` items of highest rank, one pass of 32 by default) and a FAR list with its highest rank
recorded; later scans evaluate the near list only and skip the far list while  far_max + path travelled <= max(near_max, 0)
-- the same entering row as a complete scan, hence the same iterates, flags and costs.  When the bound fails the far list is
scanned again (and the split renewed).  Reports passes of 32 items evaluated per candidate against the complete scans.
usage: scan_skip_model.py [corridors=24] [candidates per corridor and sweep=12] [near=32]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench                                    # noqa: E402
import kernel_mirror as km                      # noqa: E402
from faster_b200 import capi                    # noqa: E402
from oracle import pyoracle as po               # noqa: E402

n_corr = int(sys.argv[1]) if len(sys.argv) > 1 else 24
n_cand = int(sys.argv[2]) if len(sys.argv) > 2 else 12
NEAR = int(sys.argv[3]) if len(sys.argv) > 3 else 32
rng = np.random.default_rng(7)
w = bench.load_cfg4(0, n_corr)
pp = po.replan_pairs_port(w, threads=8)["results"]


def polys_of(kind, j):
    p, fo, Ab = w["poly_ofs_" + kind], w["face_ofs_" + kind], w["Ab_" + kind]
    return [(Ab[fo[q]:fo[q + 1], :3].copy(), Ab[fo[q]:fo[q + 1], 3].copy()) for q in range(p[j], p[j + 1])]


def item_ranks(tab, N, ff, x0, xf, dt, polys, sigma, ws):
    """ranks[k, i]: rank of item i (segment t, face f: the max over its control points, the shared one skipped when both
    segments use the same polytope -- the kernel's item list) at the iterate of scan k."""
    TZ, T0, FT = tab
    ne = 3 if ff else 2
    nz, NY = N - ne, 6 * N + 1
    Yeq = np.zeros((3, NY))
    for ax in range(3):
        s0 = np.array([x0[ax], x0[3 + ax] * dt, x0[6 + ax] * dt * dt])
        tgt = ([xf[ax]] if ff else []) + [xf[3 + ax] * dt, xf[6 + ax] * dt * dt]
        Yeq[ax] = T0[:, :3] @ s0 + T0[:, 3:] @ (np.array(tgt) - FT @ s0)
    TZN = np.sum(TZ * TZ, axis=1)
    SY = np.where(TZN > 1e-30, 1.0 / np.sqrt(np.maximum(TZN, 1e-300)), 1e15)
    out = []
    for wv in ws:
        Y = Yeq + (TZ @ wv.reshape(3, nz).T).T
        r = []
        for t in range(N):
            A, b = polys[int(sigma[t])]
            ys = [4 * N + 1 + t, 5 * N + 1 + t, t + 1] + ([t] if t == 0 or sigma[t - 1] != sigma[t] else [])
            vals = np.stack([(A @ Y[:, y] - b - km.TOL) * SY[y] for y in ys])
            r.append(vals.max(axis=0))
        out.append(np.concatenate(r))
    return np.array(out)


tot = {"candidates": 0, "scans": 0, "passes_full": 0, "passes_policy": 0, "far_rescans": 0, "later_scans": 0}
by_kind = {}
path_vs_gap = []
for kind, N, ff in (("whole", w["N_whole"], True), ("safe", w["N_safe"], False)):
    tab = capi.plan_tables(N, ff)
    k = by_kind.setdefault(kind, {"candidates": 0, "passes_full": 0, "passes_policy": 0, "later_scans": 0, "far_rescans": 0})
    for j in range(n_corr):
        x0 = w["x0"][j] if kind == "whole" else pp["R"][j]
        base = pp[kind + "_dt_base"][j]
        if not np.isfinite(base) or not np.all(np.isfinite(x0)):
            continue
        polys = polys_of(kind, j)
        fac, sig_all = w["factors_" + kind], w["sigmas_" + kind]
        for _ in range(n_cand):
            dt = fac[rng.integers(len(fac))] * base
            sigma = sig_all[rng.integers(len(sig_all))]
            trace = []
            st, *_ = km.solve(tab, N, x0, w["xf_" + kind][j], w["lim"][j], dt, polys, sigma, ff, True, True, trace=trace)
            if st < 0 or len(trace) == 0:
                continue
            R = item_ranks(tab, N, ff, x0, w["xf_" + kind][j], dt, polys, sigma, [t[0] for t in trace])
            n_items = R.shape[1]
            full_passes = -(-n_items // 32)
            near_passes = -(-min(NEAR, n_items) // 32)
            # scan 0: complete; split
            order = np.argsort(-R[0])
            near, far = order[:NEAR], order[NEAR:]
            far_max, path = (R[0][far].max() if len(far) else -np.inf), 0.0
            passes = full_passes
            for s in range(1, len(trace)):
                path += trace[s][1] * 1.0000001 + 1e-12
                m_near = R[s][near].max()
                tot["later_scans"] += 1; k["later_scans"] += 1
                if far_max + path <= max(m_near, 0.0):
                    passes += near_passes
                    assert R[s][far].max() <= max(m_near, 0.0) + 1e-9 if len(far) else True      # the bound held
                else:
                    passes += full_passes
                    tot["far_rescans"] += 1; k["far_rescans"] += 1
                    order = np.argsort(-R[s])
                    near, far = order[:NEAR], order[NEAR:]
                    far_max, path = (R[s][far].max() if len(far) else -np.inf), 0.0
            if len(trace) > 1 and len(far):
                path_vs_gap.append((sum(t[1] for t in trace[1:]), float(R[0][near].min() - R[0][far].max())))
            tot["candidates"] += 1; k["candidates"] += 1
            tot["scans"] += len(trace)
            tot["passes_full"] += full_passes * len(trace); k["passes_full"] += full_passes * len(trace)
            tot["passes_policy"] += passes; k["passes_policy"] += passes
pv = np.array(path_vs_gap)
tot["by_kind"] = by_kind
tot["near"] = NEAR
tot["passes_saved_fraction"] = 1.0 - tot["passes_policy"] / max(1, tot["passes_full"])
tot["median_total_path"] = float(np.median(pv[:, 0])) if len(pv) else None
print(json.dumps(tot))
