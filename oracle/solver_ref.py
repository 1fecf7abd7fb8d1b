"""TEST INFRASTRUCTURE ONLY -- the REFERENCE's own SolverGurobi (faster/src/solverGurobi.cpp) compiled unmodified from
/root/reference into oracle/_ref/libsolver_ref.so (oracle/solver_ref_wrap.cpp), with recording stand-ins for the Gurobi C++ API
(oracle/stub_gurobi) and for Eigen (oracle/stub_eigen).  Gives tests
  * the model the reference's code builds for one trial, row by row (`model`, `model_for_sigma`);
  * its getDTInitial, resetX / fillX (`dt_initial`, `fill_x`);
  * its genNewTraj loop end to end with an independent solver playing Gurobi's part in optimize() (`gen_new_traj`).
Only tests/ import it."""
import ctypes as C
import itertools
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libsolver_ref.so")
_lib = None
GRB_OPTIMAL, GRB_INFEASIBLE = 2, 3


def available():
    if not os.path.exists(_SO) and os.path.exists("/root/reference/faster/src/solverGurobi.cpp"):
        subprocess.call(["make", "-C", _HERE, "-s"])
    return os.path.exists(_SO)


def _L():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_SO)
        _lib.solverref_dt_initial.restype = C.c_double
        _lib.solverref_dt_initial.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    return _lib


def _f(a, n=None):
    a = np.ascontiguousarray(np.asarray(a, np.float64).reshape(-1))
    assert n is None or a.size == n
    return a


def _pack(polys):
    ofs = np.zeros(len(polys) + 1, np.int32)
    rows = []
    for p, (A, b) in enumerate(polys):
        A = np.asarray(A, float).reshape(-1, 3)
        rows.append(np.hstack([A, np.asarray(b, float).reshape(-1, 1)]))
        ofs[p + 1] = ofs[p] + len(A)
    Ab = np.ascontiguousarray(np.vstack(rows)) if rows else np.zeros((1, 4))
    return ofs, Ab


def model(N, x0, xf, lim, dt, polys, force_final=True, DC=0.01):
    """The reference's model for one trial at time allocation dt -> dict(A[n_rows, n_vars], sense (chars), rhs, ind_var, ind_val,
    vtype (chars), qdiag, n_cont = 12 N).  Variables: 12 N coefficients x[t][i] (solverGurobi.cpp:72), then (N+1) P binaries
    b[t][p] in creation order (:220-230)."""
    L = _L()
    P = len(polys)
    ofs, Ab = _pack(polys)
    nv_cap = 12 * N + (N + 1) * max(P, 1)
    nr_cap = 18 * N + 9 + 9 + 9 * (N - 1) + N + 4 * N * int(ofs[-1]) + 8
    A = np.zeros((nr_cap, nv_cap))
    sense, ind_var, ind_val = np.zeros(nr_cap, np.int32), np.zeros(nr_cap, np.int32), np.zeros(nr_cap, np.int32)
    rhs, qd = np.zeros(nr_cap), np.zeros(nv_cap)
    vt = np.zeros(nv_cap, np.uint8)
    nv, off = C.c_int(0), C.c_int(0)
    x0, xf, lim = _f(x0, 9), _f(xf, 9), _f(lim, 3)
    L.solverref_model.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_void_p,
                                  C.c_void_p] + [C.c_void_p] * 7 + [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    # the wrapper writes A with row stride n_vars: hand it a buffer and reshape afterwards
    buf = np.zeros(nr_cap * nv_cap)
    nr = L.solverref_model(N, int(bool(force_final)), x0.ctypes.data, xf.ctypes.data, lim.ctypes.data, float(DC), float(dt), P,
                           ofs.ctypes.data, Ab.ctypes.data, buf.ctypes.data, sense.ctypes.data, rhs.ctypes.data, ind_var.ctypes.data,
                           ind_val.ctypes.data, vt.ctypes.data, qd.ctypes.data, nr_cap, nv_cap, C.addressof(nv), C.addressof(off))
    assert nr >= 0, "capacity"
    n = nv.value
    A = buf[:nr * n].reshape(nr, n).copy()
    return dict(A=A, sense=np.array([chr(c) for c in sense[:nr]]), rhs=rhs[:nr].copy(), ind_var=ind_var[:nr].copy(), ind_val=ind_val[:nr].copy(),
                vtype=np.array([chr(c) for c in vt[:n]]), qdiag=qd[:n].copy(), n_cont=12 * N, offdiag=off.value, N=N, P=P)


def model_for_sigma(md, sigma):
    """The continuous QP the reference's model reduces to when the binaries select polytope sigma[t] for segment t:
    -> (qdiag[12N], Aeq, beq, Ain, bin) with Ain z <= bin, rows in the reference's creation order (box rows; per segment and face
    the four control points of the selected polytope; equalities: initial, final, continuity)."""
    N, P, nc = md["N"], md["P"], md["n_cont"]
    A, sense, rhs, iv = md["A"], md["sense"], md["rhs"], md["ind_var"]
    # the binaries alive in this trial, in creation order: b[t][p] is number t P + p of them (the reference removes and
    # re-creates them in every trial, :207-230, so their variable numbers grow from trial to trial)
    live = {int(v): k for k, v in enumerate(np.flatnonzero(md["vtype"] == "B"))}
    eq, ineq, beq, bin_ = [], [], [], []
    for k in range(len(rhs)):
        cont, bins = A[k, :nc], A[k, nc:]
        if iv[k] >= 0:                                           # indicator row on b[t][p]
            t, p = divmod(live[int(iv[k])], P)
            assert md["ind_val"][k] == 1 and sense[k] == "<" and not bins.any()
            if t < N and int(sigma[t]) == p:
                ineq.append(cont); bin_.append(rhs[k])
            continue
        if bins.any():                                           # sum_p b[t][p] == 1
            assert sense[k] == "=" and rhs[k] == 1 and not cont.any() and set(np.unique(bins)) <= {0.0, 1.0}
            continue
        if sense[k] == "=":
            eq.append(cont); beq.append(rhs[k])
        elif sense[k] == "<":
            ineq.append(cont); bin_.append(rhs[k])
        else:
            ineq.append(-cont); bin_.append(-rhs[k])
    return md["qdiag"][:nc], np.array(eq), np.array(beq), np.array(ineq), np.array(bin_)


def time_setup(N, x0, xf, lim, polys, force_final=True, DC=0.01, n_trials=20):
    """Seconds per trial of the reference's own model set-up code (no solve): a lower bound on its per-factor cost."""
    L = _L()
    ofs, Ab = _pack(polys)
    x0, xf, lim = _f(x0, 9), _f(xf, 9), _f(lim, 3)
    L.solverref_time_setup.restype = C.c_double
    L.solverref_time_setup.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    return L.solverref_time_setup(N, int(bool(force_final)), x0.ctypes.data, xf.ctypes.data, lim.ctypes.data, float(DC), len(polys),
                                  ofs.ctypes.data, Ab.ctypes.data, int(n_trials)) / n_trials


def dt_initial(x0, xf, lim, N):
    x0, xf, lim = _f(x0, 9), _f(xf, 9), _f(lim, 3)
    return float(_L().solverref_dt_initial(int(N), x0.ctypes.data, xf.ctypes.data, lim.ctypes.data))


def fill_x(N, coeffs, dt, DC):
    L = _L()
    co = _f(coeffs, 12 * N)
    cap = max(2, int(N * dt / DC) + 4)
    out = np.zeros((cap, 12))
    L.solverref_fill_x.argtypes = [C.c_int, C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_int]
    n = L.solverref_fill_x(int(N), co.ctypes.data, float(dt), float(DC), out.ctypes.data, cap)
    return out[:n].copy()


_CB = C.CFUNCTYPE(C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_int),
                  C.POINTER(C.c_int), C.POINTER(C.c_char), C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double))


def gen_new_traj(N, x0, xf, lim, polys, DC, f_init, f_final, f_inc, solve_qp, force_final=True, stop_first=False):
    """The reference's genNewTraj with `solve_qp(qdiag, Aeq, beq, Ain, bin) -> (ok, z, cost)` (any convex-QP solver) standing in
    for Gurobi: optimize() enumerates the binaries of the recorded MIQP (all P^N assignments, exactly the model's search space)
    and keeps the cheapest feasible one.  -> dict(solved, trials, dt, factor, coeffs[N,12], samples[n,12], n_optimize)."""
    L = _L()
    P = len(polys)
    ofs, Ab = _pack(polys)
    x0, xf, lim = _f(x0, 9), _f(xf, 9), _f(lim, 3)

    def cb(nv, nr, A, sense, rhs, iv, ival, vt, qd, x_out, obj_out):
        try:
            md = dict(A=np.ctypeslib.as_array(A, (nr * nv,)).reshape(nr, nv).copy(), sense=np.array([chr(sense[k]) for k in range(nr)]),
                      rhs=np.ctypeslib.as_array(rhs, (nr,)).copy(), ind_var=np.ctypeslib.as_array(iv, (nr,)).copy(),
                      ind_val=np.ctypeslib.as_array(ival, (nr,)).copy(), qdiag=np.ctypeslib.as_array(qd, (nv,)).copy(), n_cont=12 * N, N=N, P=P,
                      vtype=np.array([vt[i].decode() for i in range(nv)]))
            live_ids = np.flatnonzero(md["vtype"] == "B")
            best = None
            for sigma in (itertools.product(range(P), repeat=N) if P else [()]):
                q, Aeq, beq, Ain, bin_ = model_for_sigma(md, sigma) if P else model_for_sigma(dict(md, P=1), [0] * N)
                ok, z, cost = solve_qp(q, Aeq, beq, Ain, bin_)
                if ok and (best is None or cost < best[0]):
                    best = (cost, z, sigma)
            if best is None:
                return GRB_INFEASIBLE
            for i in range(12 * N):
                x_out[i] = best[1][i]
            for t in range(N):                                   # the binaries of the winning assignment (row N stays 0)
                if P:
                    x_out[int(live_ids[t * P + best[2][t]])] = 1.0
            obj_out[0] = best[0]
            return GRB_OPTIMAL
        except Exception as e:                                   # never raise through C
            print("solver_ref callback failed:", repr(e))
            return 12                                            # GRB_NUMERIC
    cfn = _CB(cb)
    trials, ns, nopt = C.c_int(0), C.c_int(0), C.c_int(0)
    dt, fac = C.c_double(0), C.c_double(0)
    co = np.zeros((N, 12))
    cap = 200000
    samples = np.zeros((cap, 12))
    L.solverref_gen_new_traj.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_void_p, C.c_void_p,
                                         C.c_double, C.c_double, C.c_double, _CB, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    solved = L.solverref_gen_new_traj(N, int(bool(force_final)), x0.ctypes.data, xf.ctypes.data, lim.ctypes.data, float(DC), P, ofs.ctypes.data,
                                      Ab.ctypes.data, float(f_init), float(f_final), float(f_inc), cfn, int(bool(stop_first)),
                                      C.addressof(trials), C.addressof(dt), C.addressof(fac), co.ctypes.data, samples.ctypes.data, cap,
                                      C.addressof(ns), C.addressof(nopt))
    return dict(solved=bool(solved), trials=trials.value, dt=dt.value, factor=fac.value, coeffs=co, samples=samples[:ns.value].copy(),
                n_optimize=nopt.value)
