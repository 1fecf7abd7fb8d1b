#!/usr/bin/env python
"""List-scheduling model of the persistent kernel's claim order (CPU only; evidence for DESIGN.md section 3/9).

Per-candidate solve lengths = active-set changes counted by the numpy mirror of the specialised kernel
(tests/kernel_mirror.py, normalised pivoting + thin factorisation) on bench.py's workload.  The model replays the
kernel's scheduling -- CTAs of 4 warps adopt problems cyclically from a CTA-specific start, warps claim candidates one
by one, a CTA changes problem behind a block barrier -- for different claim orders and prints the busy fraction
(mean warp work / makespan).  usage: claim_order_model.py [corridors]   (default 16; ~15 s per 16 on 8 cores)"""
import heapq
import os
import sys
from multiprocessing import Pool

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench                      # noqa: E402
import kernel_mirror as km        # noqa: E402
from faster_b200 import capi      # noqa: E402

NC = int(sys.argv[1]) if len(sys.argv) > 1 else 16
WORK = {"whole": bench.make_workload(NC, 0, "whole"), "safe": bench.make_workload(NC, 100000, "safe")}


def lengths(job):
    kind, c = job
    w = WORK[kind]
    tabs = capi.plan_tables(w["N"], w["ff"])
    pb = w["probs"][c]
    out = np.zeros(bench.CAND, np.int32)
    for i in range(bench.CAND):
        k = c * bench.CAND + i
        out[i] = km.solve(tabs, w["N"], pb["x0"], pb["xf"], pb["lim"], w["dt"][k], pb["polys"], w["sigma"][k], w["ff"],
                          normalised=True, thin=True)[3]
    return kind, c, out


def simulate(its, n_cta, warps=4, setup=1.0, switch=0.3):
    """its[problem, k] = length of the k-th claimed candidate.  Returns (makespan, mean work per warp)."""
    n_prob, count = its.shape
    counters = [0] * n_prob
    cursor = [(b * n_prob) // n_cta for b in range(n_cta)]
    visited = [0] * n_cta
    prob = [None] * n_cta
    waiting = [0] * n_cta
    latest = [0.0] * n_cta
    busy = t_end = 0.0

    def adopt(b):
        while visited[b] < n_prob:
            pj = (cursor[b] + visited[b]) % n_prob
            visited[b] += 1
            if counters[pj] < count:
                return pj
        return None

    ev = []
    for b in range(n_cta):
        prob[b] = adopt(b)
        if prob[b] is not None:
            for w in range(warps):
                heapq.heappush(ev, (0.0, b, w))
    while ev:
        t, b, w = heapq.heappop(ev)
        p = prob[b]
        if p is not None and counters[p] < count:
            d = setup + its[p, counters[p]]
            counters[p] += 1
            busy += d
            t_end = max(t_end, t + d)
            heapq.heappush(ev, (t + d, b, w))
        else:                                   # block barrier before the CTA adopts its next problem
            waiting[b] += 1
            latest[b] = max(latest[b], t)
            if waiting[b] == warps:
                waiting[b] = 0
                prob[b] = adopt(b)
                if prob[b] is not None:
                    for ww in range(warps):
                        heapq.heappush(ev, (latest[b] + switch, b, ww))
    return t_end, busy / (n_cta * warps)


if __name__ == "__main__":
    with Pool(min(8, os.cpu_count() or 1)) as pool:
        res = pool.map(lengths, [(k, c) for k in ("whole", "safe") for c in range(NC)])
    L = {"whole": np.zeros((NC, bench.CAND)), "safe": np.zeros((NC, bench.CAND))}
    for kind, c, o in res:
        L[kind][c] = o
    n_cta = int(round(NC * 592 / 64))           # the bench's ratio: 592 CTAs for 64 corridors
    for kind in ("whole", "safe"):
        its = L[kind]
        grid = its.reshape(NC, bench.N_DT, bench.N_SIG)
        print("%s: mean %.2f  p99 %.0f  max %d changes;  by time allocation: mean %s  max %s" %
              (kind, its.mean(), np.percentile(its, 99), its.max(), np.round(grid.mean((0, 2)), 1).tolist(),
               grid.max((0, 2)).astype(int).tolist()))
        adj = [np.corrcoef(grid[:, d].ravel(), grid[:, d + 1].ravel())[0, 1] for d in range(bench.N_DT - 1)]
        print("   correlation of lengths between adjacent time allocations (same assignment): %s" % np.round(adj, 2).tolist())
        orders = {"ascending (first candidate first)": its, "descending (product)": its[:, ::-1],
                  "longest first within a problem (needs the lengths)": -np.sort(-its, axis=1)}
        for name, arr in orders.items():
            te, avg = simulate(arr, n_cta)
            te1, avg1 = simulate(arr, n_cta * 4, warps=1)
            print("   %-52s busy %.1f %%   (warps adopting problems independently: %.1f %%)" % (name, 100 * avg / te, 100 * avg1 / te1))
