"""TEST INFRASTRUCTURE ONLY -- ctypes loader for oracle/_build/libfq_oracle.so (the CPU restatement).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libfq_oracle.so")
_lib = None

_d = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_i = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_u8 = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")


_PORT_SO = os.path.join(_HERE, "_build", "libfq_cpu_port.so")
_port = None


def build(force=False):
    src = os.path.join(_HERE, "fq_oracle.c")
    src2 = os.path.join(_HERE, "fq_cpu_port.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src) or \
            not os.path.exists(_PORT_SO) or os.path.getmtime(_PORT_SO) < os.path.getmtime(src2):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        L.fqo_dt_initial.restype = C.c_double
        L.fqo_dt_initial.argtypes = [_d, _d, _d, C.c_int]
        L.fqo_num_samples.restype = C.c_int
        L.fqo_num_samples.argtypes = [C.c_int, C.c_double, C.c_double]
        L.fqo_fill_x.restype = None
        L.fqo_fill_x.argtypes = [C.c_int, _d, C.c_double, C.c_double, C.c_int, _d]
        L.fqo_set_row_tol.restype = None
        L.fqo_set_row_tol.argtypes = [C.c_double]
        L.fqo_get_row_tol.restype = C.c_double
        _lib = L
    return _lib


def set_row_tol(tol):
    """Row tolerance of every later solve (mirrors the product's fq_set_option("row_tol_1e9", ...))."""
    lib().fqo_set_row_tol(float(tol))


def get_row_tol():
    return lib().fqo_get_row_tol()


def pack_polys(polys):
    """list of (A[F,3], b[F]) -> (P, face_ofs int32[P+1], Ab float64[SF,4])."""
    ofs = [0]
    rows = []
    for A, b in polys:
        A = np.asarray(A, float).reshape(-1, 3)
        b = np.asarray(b, float).reshape(-1)
        rows.append(np.hstack([A, b[:, None]]))
        ofs.append(ofs[-1] + A.shape[0])
    Ab = np.ascontiguousarray(np.vstack(rows) if rows else np.zeros((0, 4)), dtype=np.float64)
    if Ab.shape[0] == 0:
        Ab = np.zeros((1, 4))
    return len(polys), np.asarray(ofs, np.int32), Ab


def _a(x, n):
    x = np.ascontiguousarray(np.asarray(x, np.float64).reshape(-1))
    assert x.size == n
    return x


def solve_fixed(N, x0, xf, lim, dt, polys, sigma, force_final=True):
    """-> (status, cost, coeffs[N,12], iters); status 1 optimal, 0 infeasible, -1 numeric."""
    L = lib()
    P, ofs, Ab = pack_polys(polys)
    sig = np.ascontiguousarray(np.asarray(sigma if sigma is not None else np.zeros(N), np.uint8))
    cost = C.c_double(np.inf)
    it = C.c_int(0)
    co = np.zeros(12 * N)
    rc = L.fqo_solve_fixed(C.c_int(N), C.c_int(int(force_final)), _a(x0, 9).ctypes, _a(xf, 9).ctypes,
                           _a(lim, 3).ctypes, C.c_int(P), ofs.ctypes, Ab.ctypes, C.c_double(dt), sig.ctypes,
                           C.byref(cost), co.ctypes, C.byref(it))
    return rc, cost.value, co.reshape(N, 12), it.value


def solve_batch(N, x0, xf, lim, polys, dts, sigmas, force_final=True, want_coeffs=False, threads=1):
    L = lib()
    P, ofs, Ab = pack_polys(polys)
    dts = np.ascontiguousarray(dts, np.float64)
    n = dts.size
    sig = np.ascontiguousarray(np.asarray(sigmas, np.uint8).reshape(n, N))
    feas = np.zeros(n, np.uint8)
    cost = np.zeros(n)
    co = np.zeros((n, N, 12)) if want_coeffs else None
    L.fqo_solve_batch(C.c_int(N), C.c_int(int(force_final)), _a(x0, 9).ctypes, _a(xf, 9).ctypes, _a(lim, 3).ctypes,
                      C.c_int(P), ofs.ctypes, Ab.ctypes, C.c_int(n), dts.ctypes, sig.ctypes, feas.ctypes,
                      cost.ctypes, co.ctypes if want_coeffs else None, C.c_int(threads))
    return feas, cost, co


def solve_multi(N, force_final, x0, xf, lim, poly_ofs, face_ofs, Ab, cand_ofs, dts, sigmas, threads=1):
    """Same array layout as the product's fq_solve_multi (host arrays).  -> (feasible, cost)."""
    L = lib()
    n_prob = len(cand_ofs) - 1
    n = int(cand_ofs[-1])
    c = lambda a, t: np.ascontiguousarray(a, t)
    x0, xf, lim, Ab, dts = c(x0, np.float64), c(xf, np.float64), c(lim, np.float64), c(Ab, np.float64), c(dts, np.float64)
    poly_ofs, face_ofs, cand_ofs = c(poly_ofs, np.int32), c(face_ofs, np.int32), c(cand_ofs, np.int32)
    sig = c(sigmas, np.uint8)
    feas = np.zeros(n, np.uint8)
    cost = np.zeros(n)
    L.fqo_solve_multi(C.c_int(N), C.c_int(int(force_final)), C.c_int(n_prob), x0.ctypes, xf.ctypes, lim.ctypes,
                      poly_ofs.ctypes, face_ofs.ctypes, Ab.ctypes, cand_ofs.ctypes, dts.ctypes, sig.ctypes,
                      feas.ctypes, cost.ctypes, C.c_int(threads))
    return feas, cost


def solve_miqp(N, x0, xf, lim, dt, polys, force_final=True):
    """Branch and bound over ALL sigma in P^N -> (status, cost, coeffs, sigma, nodes)."""
    L = lib()
    P, ofs, Ab = pack_polys(polys)
    sig = np.zeros(N, np.uint8)
    cost = C.c_double(np.inf)
    nodes = C.c_long(0)
    co = np.zeros(12 * N)
    rc = L.fqo_solve_miqp(C.c_int(N), C.c_int(int(force_final)), _a(x0, 9).ctypes, _a(xf, 9).ctypes,
                          _a(lim, 3).ctypes, C.c_int(P), ofs.ctypes, Ab.ctypes, C.c_double(dt), sig.ctypes,
                          C.byref(cost), co.ctypes, C.byref(nodes))
    return rc, cost.value, co.reshape(N, 12), sig, nodes.value


def dt_initial(x0, xf, lim, N):
    return lib().fqo_dt_initial(_a(x0, 9), _a(xf, 9), _a(lim, 3), N)


def gen_new_traj(N, x0, xf, lim, polys, DC, f_init, f_final, f_inc, sigma_list=None, force_final=True):
    """-> dict(solved, dt, factor, trials, sigma, cost, coeffs)."""
    L = lib()
    P, ofs, Ab = pack_polys(polys)
    if sigma_list is not None:
        sl = np.ascontiguousarray(np.asarray(sigma_list, np.uint8).reshape(-1, N))
        ns, slp = sl.shape[0], sl.ctypes
    else:
        ns, slp = 0, None
    dt = C.c_double(0)
    fac = C.c_double(0)
    tr = C.c_int(0)
    cost = C.c_double(np.inf)
    sig = np.zeros(N, np.uint8)
    co = np.zeros(12 * N)
    rc = L.fqo_gen_new_traj(C.c_int(N), C.c_int(int(force_final)), _a(x0, 9).ctypes, _a(xf, 9).ctypes,
                            _a(lim, 3).ctypes, C.c_int(P), ofs.ctypes, Ab.ctypes, C.c_double(DC),
                            C.c_double(f_init), C.c_double(f_final), C.c_double(f_inc), C.c_int(ns), slp,
                            C.byref(dt), C.byref(fac), C.byref(tr), sig.ctypes, C.byref(cost), co.ctypes)
    return dict(solved=bool(rc), dt=dt.value, factor=fac.value, trials=tr.value, sigma=sig, cost=cost.value,
                coeffs=co.reshape(N, 12))


def fill_x(N, coeffs, dt, DC):
    L = lib()
    n = L.fqo_num_samples(N, dt, DC)
    out = np.zeros((n, 12))
    L.fqo_fill_x(N, np.ascontiguousarray(coeffs, np.float64).reshape(-1), dt, DC, n, out.reshape(-1))
    return out


def port_lib():
    """oracle/fq_cpu_port.c: the tuned CPU port of the GPU kernel's algorithm (bench.py's CPU arm)."""
    global _port
    if _port is None:
        if not os.path.exists(_PORT_SO):
            build()
        _port = C.CDLL(_PORT_SO)
        _port.fqc_set_row_tol.argtypes = [C.c_double]
        _port.fqc_set_row_tol.restype = None
    return _port


def solve_multi_port(N, force_final, x0, xf, lim, poly_ofs, face_ofs, Ab, cand_ofs, dts, sigmas, threads=1, want_coeffs=False,
                     want_iters=False):
    """fq_solve_multi's layout through the tuned CPU port.  -> (feasible, cost[, coeffs][, iters])."""
    L = port_lib()
    n_prob = len(cand_ofs) - 1
    n = int(cand_ofs[-1])
    c = lambda a, t: np.ascontiguousarray(a, t)
    x0, xf, lim, Ab, dts = c(x0, np.float64), c(xf, np.float64), c(lim, np.float64), c(Ab, np.float64), c(dts, np.float64)
    poly_ofs, face_ofs, cand_ofs = c(poly_ofs, np.int32), c(face_ofs, np.int32), c(cand_ofs, np.int32)
    sig = c(sigmas, np.uint8)
    feas = np.zeros(n, np.uint8)
    cost = np.zeros(n)
    co = np.zeros((n, N, 12)) if want_coeffs else None
    it = np.zeros(n, np.int32) if want_iters else None
    rc = L.fqc_solve_multi(C.c_int(N), C.c_int(int(force_final)), C.c_int(n_prob), x0.ctypes, xf.ctypes, lim.ctypes, poly_ofs.ctypes,
                           face_ofs.ctypes, Ab.ctypes, cand_ofs.ctypes, dts.ctypes, sig.ctypes, feas.ctypes, cost.ctypes,
                           co.ctypes if want_coeffs else None, it.ctypes if want_iters else None, C.c_int(threads))
    if rc != 0:
        raise RuntimeError("fqc_solve_multi failed")
    out = (feas, cost)
    if want_coeffs:
        out += (co,)
    if want_iters:
        out += (it,)
    return out


def replan_pairs_port(w, threads=1, want_coeffs=True):
    """The whole chained replan in C (fqc_replan_pairs of fq_cpu_port.c) on a pair-workload dict.  -> dict like
    pair_oracle.replan_pairs (results as a structured array with fq_pair_result's fields)."""
    L = port_lib()
    n, Nw, Ns = w["n_prob"], w["N_whole"], w["N_safe"]
    c = lambda a, t: np.ascontiguousarray(a, t)
    fw, fs = c(w["factors_whole"], np.float64), c(w["factors_safe"], np.float64)
    sw, ss = c(w["sigmas_whole"], np.uint8), c(w["sigmas_safe"], np.uint8)
    ncw, ncs = n * len(fw) * len(sw), n * len(fs) * len(ss)
    dt = np.dtype([("whole_dt_index", np.int32), ("whole_sigma_index", np.int32), ("safe_dt_index", np.int32),
                   ("safe_sigma_index", np.int32), ("whole_cost", np.float64), ("safe_cost", np.float64), ("whole_dt", np.float64),
                   ("safe_dt", np.float64), ("whole_dt_base", np.float64), ("safe_dt_base", np.float64),
                   ("n_samples_whole", np.int32), ("k_safe", np.int32), ("R", np.float64, (9,))])
    out = dict(results=np.zeros(n, dt), feasible_whole=np.zeros(ncw, np.uint8), cost_whole=np.zeros(ncw),
               feasible_safe=np.zeros(ncs, np.uint8), cost_safe=np.zeros(ncs), coeffs_whole=np.zeros((n, Nw, 12)),
               coeffs_safe=np.zeros((n, Ns, 12)) if want_coeffs else None)
    a = [c(w[k], np.float64) for k in ("x0", "xf_whole", "xf_safe", "lim")]
    pw, fow, Aw = c(w["poly_ofs_whole"], np.int32), c(w["face_ofs_whole"], np.int32), c(w["Ab_whole"], np.float64)
    ps, fos, As = c(w["poly_ofs_safe"], np.int32), c(w["face_ofs_safe"], np.int32), c(w["Ab_safe"], np.float64)
    rc = L.fqc_replan_pairs(C.c_int(n), C.c_int(Nw), C.c_int(Ns), C.c_double(w["DC"]), C.c_double(w["r_fraction"]), a[0].ctypes,
                            a[1].ctypes, a[2].ctypes, a[3].ctypes, pw.ctypes, fow.ctypes, Aw.ctypes, ps.ctypes, fos.ctypes, As.ctypes,
                            C.c_int(len(fw)), fw.ctypes, C.c_int(len(sw)), sw.ctypes, C.c_int(len(fs)), fs.ctypes, C.c_int(len(ss)),
                            ss.ctypes, out["feasible_whole"].ctypes, out["cost_whole"].ctypes, out["feasible_safe"].ctypes,
                            out["cost_safe"].ctypes, out["coeffs_whole"].ctypes,
                            out["coeffs_safe"].ctypes if want_coeffs else None, out["results"].ctypes, C.c_int(threads))
    if rc != 0:
        raise RuntimeError("fqc_replan_pairs failed")
    return out
