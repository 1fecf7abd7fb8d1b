"""TEST INFRASTRUCTURE ONLY -- the chained replan (whole sweep -> R -> safe sweep) restated with the CPU oracle.

Follows Faster::replan(): faster/src/faster.cpp:406-430 (whole: setX0(A), setXf(E), setPolytopes, genNewTraj, fillX),
:474-475 (R = X_temp_[k_safe]), :521-537 (safe: setX0(R), setXf(M), setForceFinalConstraint(false), genNewTraj), with
genNewTraj's sweep as solverGurobi.cpp:445-472 (ascending factors, first feasible wins, dt = factor * max(getDTInitial,
2 DC): :494-497) restricted to the supplied assignment list, and k_safe = min(n - 1, (int)(r_fraction n)) standing in for
findIndexR (faster.cpp:173-216).  Only tests/ and bench.py's checker / CPU legs import this.
"""
import numpy as np

from . import pyoracle as po


def _expand(n_prob, factors, sigmas, dt_base):
    nf, ns = len(factors), len(sigmas)
    dts = (dt_base[:, None, None] * np.asarray(factors)[None, :, None] * np.ones((1, 1, ns))).reshape(-1)
    sig = np.broadcast_to(sigmas[None, None], (n_prob, nf, ns, sigmas.shape[1])).reshape(-1, sigmas.shape[1])
    cand_ofs = (np.arange(n_prob + 1) * nf * ns).astype(np.int32)
    return np.ascontiguousarray(dts), np.ascontiguousarray(sig), cand_ofs


def _select(feas, cost, n_prob, nf, ns):
    """first feasible factor, then min cost, then lowest index -> (dt index, sigma index) per problem (-1: none)."""
    f = feas.reshape(n_prob, nf, ns).astype(bool)
    c = cost.reshape(n_prob, nf, ns)
    di = np.full(n_prob, -1, np.int32)
    si = np.full(n_prob, -1, np.int32)
    for j in range(n_prob):
        rows = np.flatnonzero(f[j].any(axis=1))
        if len(rows):
            d = rows[0]
            cc = np.where(f[j, d], c[j, d], np.inf)
            di[j], si[j] = d, int(np.argmin(cc))
    return di, si


def _sub_polys(po_, fo_, Ab, j):
    return [(Ab[fo_[p]:fo_[p + 1], :3], Ab[fo_[p]:fo_[p + 1], 3]) for p in range(po_[j], po_[j + 1])]


def replan_pairs(w, threads=1, dt_base_whole=None, dt_base_safe=None, fast=False):
    """w: dict as faster_b200.capi.make_pair_workload builds.  dt_base_* (optional): use these time-allocation bases (e.g.
    the ones the device computed) instead of the oracle's own getDTInitial, so that a solve comparison is not disturbed
    by a last-bit difference in dt.  -> dict with the fields of fq_pair_result as arrays, per-candidate flags / costs,
    and the winners' coefficients.  fast=True: the sweeps run through the tuned CPU port (oracle/fq_cpu_port.c) instead of
    the literal restatement (fq_oracle.c): bench.py's CPU arm."""
    solve_multi = (lambda *a: po.solve_multi_port(*a)[:2]) if fast else po.solve_multi
    n, Nw, Ns, DC = w["n_prob"], w["N_whole"], w["N_safe"], w["DC"]
    fw, fs, sw, ss = w["factors_whole"], w["factors_safe"], w["sigmas_whole"], w["sigmas_safe"]
    out = {}
    dbw = np.array([max(po.dt_initial(w["x0"][j], w["xf_whole"][j], w["lim"][j], Nw), 2 * DC) for j in range(n)])
    out["whole_dt_base_own"] = dbw.copy()
    if dt_base_whole is not None:
        dbw = np.asarray(dt_base_whole, float)
    dts, sig, co = _expand(n, fw, sw, dbw)
    feas_w, cost_w = solve_multi(Nw, True, w["x0"], w["xf_whole"], w["lim"], w["poly_ofs_whole"], w["face_ofs_whole"],
                                    w["Ab_whole"], co, dts, sig, threads)
    di, si = _select(feas_w, cost_w, n, len(fw), len(sw))
    R = np.full((n, 9), np.nan)
    nsamp = np.zeros(n, np.int32)
    ksafe = np.full(n, -1, np.int32)
    coeffs_w = np.zeros((n, Nw, 12))
    wcost = np.full(n, np.inf)
    wdt = np.full(n, np.nan)
    if fast and (di >= 0).any():
        # winners' coefficients in one batch through the port (one candidate per corridor)
        dtw = np.where(di >= 0, np.asarray(fw)[np.maximum(di, 0)] * dbw, 1.0)
        fwn, cwn, cfw = po.solve_multi_port(Nw, True, w["x0"], w["xf_whole"], w["lim"], w["poly_ofs_whole"], w["face_ofs_whole"],
                                            w["Ab_whole"], np.arange(n + 1, dtype=np.int32), dtw, sw[np.maximum(si, 0)], threads,
                                            want_coeffs=True)
    for j in range(n):
        if di[j] < 0:
            continue
        dt = fw[di[j]] * dbw[j]
        if fast:
            st, c, cf = int(fwn[j]), cwn[j], cfw[j]
        else:
            st, c, cf, _ = po.solve_fixed(Nw, w["x0"][j], w["xf_whole"][j], w["lim"][j], dt,
                                          _sub_polys(w["poly_ofs_whole"], w["face_ofs_whole"], w["Ab_whole"], j), sw[si[j]], True)
        assert st == 1
        coeffs_w[j], wcost[j], wdt[j] = cf, c, dt
        X = po.fill_x(Nw, cf, dt, DC)
        k = min(len(X) - 1, int(w["r_fraction"] * len(X)))
        R[j], nsamp[j], ksafe[j] = X[k, :9], len(X), k
    have = di >= 0
    dbs = np.full(n, np.nan)
    for j in np.flatnonzero(have):
        dbs[j] = max(po.dt_initial(R[j], w["xf_safe"][j], w["lim"][j], Ns), 2 * DC)
    out["safe_dt_base_own"] = dbs.copy()
    if dt_base_safe is not None:
        dbs = np.where(have, np.asarray(dt_base_safe, float), np.nan)
    nfs, nss = len(fs), len(ss)
    feas_s = np.zeros(n * nfs * nss, np.uint8)
    cost_s = np.full(n * nfs * nss, np.inf)
    coeffs_s = np.zeros((n, Ns, 12))
    sdi = np.full(n, -1, np.int32)
    ssi = np.full(n, -1, np.int32)
    scost = np.full(n, np.inf)
    sdt = np.full(n, np.nan)
    idx = np.flatnonzero(have)
    if len(idx):
        # the sub-batch of corridors that have a whole trajectory
        po_s, fo_s = w["poly_ofs_safe"], w["face_ofs_safe"]
        sub_po, sub_fo, rows = [0], [0], []
        for j in idx:
            for p in range(po_s[j], po_s[j + 1]):
                rows.append(w["Ab_safe"][fo_s[p]:fo_s[p + 1]])
                sub_fo.append(sub_fo[-1] + fo_s[p + 1] - fo_s[p])
            sub_po.append(sub_po[-1] + po_s[j + 1] - po_s[j])
        Ab = np.ascontiguousarray(np.vstack(rows)) if rows else np.zeros((1, 4))
        dts, sig, co = _expand(len(idx), fs, ss, dbs[idx])
        f, c = solve_multi(Ns, False, np.ascontiguousarray(R[idx]), np.ascontiguousarray(w["xf_safe"][idx]),
                              np.ascontiguousarray(w["lim"][idx]), np.array(sub_po, np.int32), np.array(sub_fo, np.int32), Ab,
                              co, dts, sig, threads)
        per = nfs * nss
        for q, j in enumerate(idx):
            feas_s[j * per:(j + 1) * per] = f[q * per:(q + 1) * per]
            cost_s[j * per:(j + 1) * per] = c[q * per:(q + 1) * per]
        d2, s2 = _select(f, c, len(idx), nfs, nss)
        if fast and (d2 >= 0).any():
            dtq = np.where(d2 >= 0, np.asarray(fs)[np.maximum(d2, 0)] * dbs[idx], 1.0)
            fsn, csn, cfs = po.solve_multi_port(Ns, False, np.ascontiguousarray(R[idx]), np.ascontiguousarray(w["xf_safe"][idx]),
                                                np.ascontiguousarray(w["lim"][idx]), np.array(sub_po, np.int32), np.array(sub_fo, np.int32),
                                                Ab, np.arange(len(idx) + 1, dtype=np.int32), dtq, ss[np.maximum(s2, 0)], threads,
                                                want_coeffs=True)
        for q, j in enumerate(idx):
            if d2[q] < 0:
                continue
            sdi[j], ssi[j] = d2[q], s2[q]
            dt = fs[d2[q]] * dbs[j]
            if fast:
                st, cc, cf = int(fsn[q]), csn[q], cfs[q]
            else:
                st, cc, cf, _ = po.solve_fixed(Ns, R[j], w["xf_safe"][j], w["lim"][j], dt, _sub_polys(po_s, fo_s, w["Ab_safe"], j),
                                               ss[s2[q]], False)
            assert st == 1
            coeffs_s[j], scost[j], sdt[j] = cf, cc, dt
    out.update(whole_dt_index=di, whole_sigma_index=si, safe_dt_index=sdi, safe_sigma_index=ssi, whole_cost=wcost,
               safe_cost=scost, whole_dt=wdt, safe_dt=sdt, whole_dt_base=dbw, safe_dt_base=dbs, n_samples_whole=nsamp,
               k_safe=ksafe, R=R, feasible_whole=feas_w, cost_whole=cost_w, feasible_safe=feas_s, cost_safe=cost_s,
               coeffs_whole=coeffs_w, coeffs_safe=coeffs_s)
    return out
