"""TEST INFRASTRUCTURE ONLY -- literal full-space restatement of the SolverGurobi model.

PARITY UNPINNED: the reference delegates the arithmetic to Gurobi (closed source, absent from
/root/reference and from this image) and its tree holds no golden outputs for this
path.  This file rebuilds the model *exactly as the reference writes it* -- in the reference's own 12N
coefficient variables, with every equality and inequality row spelled out, no elimination and no
condensing -- and hands it to an independent solver (HiGHS, vendored in scipy) for one fixed
interval->polytope assignment sigma.  It is the "second opinion" that pins oracle/fq_oracle.c and, through
it, the CUDA path.  Only tests/ may import it.

What IS pinned to the reference: `build` is compared, row by row and coefficient by coefficient (exact equality), with the
model that the reference's own solverGurobi.cpp creates when it is compiled from /root/reference against a recording
stand-in for the Gurobi API (oracle/solver_ref.py, tests/test_reference_solver_cpu.py).  What stays unpinned is Gurobi's
numerical solve of that model (tolerances, MIP search).

Variable order (solverGurobi.cpp:72):  z[12*t + k],  k = ax ay az bx by bz cx cy cz dx dy dz,
p_t(tau) = a tau^3 + b tau^2 + c tau + d  (solverGurobi.cpp:761-788).
"""
import numpy as np
import scipy.sparse as sp

INF = np.inf


def _pos(N, t, tau, ax):            # solverGurobi.cpp:761-767
    r = np.zeros(12 * N)
    r[12 * t + 0 + ax] = tau * tau * tau     # products in the order the reference writes them: x * tau * tau * tau
    r[12 * t + 3 + ax] = tau * tau
    r[12 * t + 6 + ax] = tau
    r[12 * t + 9 + ax] = 1.0
    return r


def _vel(N, t, tau, ax):            # solverGurobi.cpp:769-774
    r = np.zeros(12 * N)
    r[12 * t + 0 + ax] = 3 * tau * tau           # 3 * x * tau * tau
    r[12 * t + 3 + ax] = 2 * tau
    r[12 * t + 6 + ax] = 1.0
    return r


def _acc(N, t, tau, ax):            # solverGurobi.cpp:776-781
    r = np.zeros(12 * N)
    r[12 * t + 0 + ax] = 6 * tau
    r[12 * t + 3 + ax] = 2.0
    return r


def _jerk(N, t, ax):                # solverGurobi.cpp:783-788
    r = np.zeros(12 * N)
    r[12 * t + 0 + ax] = 6.0
    return r


def _cp(N, t, k, dt, ax):
    """Bezier control point k of segment t, axis ax (solverGurobi.cpp:812-862)."""
    if k == 0:                                   # getCP0: getPos(t, 0)
        return _pos(N, t, 0.0, ax)
    if k == 3:                                   # getCP3: getPos(t, dt)
        return _pos(N, t, dt, ax)
    r = np.zeros(12 * N)
    cn, dn, bn = dt, 1.0, dt * dt               # getCn/getDn/getBn scale factors (:817-830)
    if k == 1:                                   # (Cn + 3 Dn)/3            (:840-847)
        r[12 * t + 6 + ax] = cn / 3.0
        r[12 * t + 9 + ax] = 3.0 * dn / 3.0
    else:                                        # (Bn + 2 Cn + 3 Dn)/3     (:849-856)
        r[12 * t + 3 + ax] = bn / 3.0
        r[12 * t + 6 + ax] = 2.0 * cn / 3.0
        r[12 * t + 9 + ax] = 3.0 * dn / 3.0
    return r


def build(N, x0, xf, lim, dt, polys, sigma, force_final=True):
    """Rows of the model for a FIXED assignment sigma (len N, polytope index per segment).

    polys: list of (A [F,3], b [F]).  Returns (Q, Aeq, beq, Ain, bin) with cost 0.5 z'Qz,
    Aeq z = beq, Ain z <= bin.  Q is the Hessian of sum_t sum_ax (6 a)^2 (solverGurobi.cpp:113-119).
    """
    x0 = np.asarray(x0, float)
    xf = np.asarray(xf, float)
    n = 12 * N
    q = np.zeros(n)
    for t in range(N):
        for ax in range(3):
            q[12 * t + ax] = 2.0 * 36.0          # d2/da2 of (6a)^2
    Q = sp.diags(q).tocsc()
    Aeq, beq, Ain, bin_ = [], [], [], []
    for ax in range(3):                          # setConstraintsX0 (:359-380)
        Aeq.append(_pos(N, 0, 0.0, ax)); beq.append(x0[ax])
        Aeq.append(_vel(N, 0, 0.0, ax)); beq.append(x0[3 + ax])
        Aeq.append(_acc(N, 0, 0.0, ax)); beq.append(x0[6 + ax])
    for ax in range(3):                          # setConstraintsXf (:332-357)
        if force_final:
            Aeq.append(_pos(N, N - 1, dt, ax)); beq.append(xf[ax])
        Aeq.append(_vel(N, N - 1, dt, ax)); beq.append(xf[3 + ax])
        Aeq.append(_acc(N, N - 1, dt, ax)); beq.append(xf[6 + ax])
    for t in range(N - 1):                       # setDynamicConstraints (:499-524)
        for ax in range(3):
            Aeq.append(_pos(N, t, dt, ax) - _pos(N, t + 1, 0.0, ax)); beq.append(0.0)
            Aeq.append(_vel(N, t, dt, ax) - _vel(N, t + 1, 0.0, ax)); beq.append(0.0)
            Aeq.append(_acc(N, t, dt, ax) - _acc(N, t + 1, 0.0, ax)); beq.append(0.0)
    for t in range(N):                           # setMaxConstraints (:390-407)
        for ax in range(3):
            for row, m in ((_vel(N, t, 0.0, ax), lim[0]), (_acc(N, t, 0.0, ax), lim[1]),
                           (_jerk(N, t, ax), lim[2])):
                Ain.append(row); bin_.append(m)
                Ain.append(-row); bin_.append(m)
    if len(polys) > 0:                           # setPolytopesConstraints (:217-290), b[t][sigma[t]] = 1
        for t in range(N):
            A, b = polys[int(sigma[t])]
            A = np.asarray(A, float)
            for k in range(4):
                cps = [_cp(N, t, k, dt, ax) for ax in range(3)]
                for f in range(A.shape[0]):
                    Ain.append(A[f, 0] * cps[0] + A[f, 1] * cps[1] + A[f, 2] * cps[2])
                    bin_.append(float(b[f]))
    return (Q, np.array(Aeq), np.array(beq), np.array(Ain), np.array(bin_))


def solve_qp_highs(Q, Aeq, beq, Ain, bin_, with_status=False):
    """min 0.5 z'Qz  s.t.  Aeq z = beq, Ain z <= bin  by HiGHS (scipy-vendored).  -> (feasible, z[, status name])."""
    from scipy.optimize._highspy import _core as h
    n = Q.shape[0]
    A = sp.csc_matrix(np.vstack([Aeq, Ain]) if len(bin_) else np.asarray(Aeq))
    lo = np.concatenate([beq, np.full(len(bin_), -INF)])
    up = np.concatenate([beq, bin_])
    model = h.HighsModel()
    lp = model.lp_
    lp.num_col_ = n
    lp.num_row_ = A.shape[0]
    lp.col_cost_ = np.zeros(n)
    lp.col_lower_ = np.full(n, -INF)
    lp.col_upper_ = np.full(n, INF)
    lp.row_lower_ = lo
    lp.row_upper_ = up
    lp.a_matrix_.format_ = h.MatrixFormat.kColwise
    lp.a_matrix_.start_ = A.indptr.astype(np.int32)
    lp.a_matrix_.index_ = A.indices.astype(np.int32)
    lp.a_matrix_.value_ = A.data.astype(float)
    Qt = sp.csc_matrix(sp.tril(sp.coo_matrix(Q)))
    hs = model.hessian_
    hs.dim_ = n
    hs.format_ = h.HessianFormat.kTriangular
    hs.start_ = Qt.indptr.astype(np.int32)
    hs.index_ = Qt.indices.astype(np.int32)
    hs.value_ = Qt.data.astype(float)
    H = h._Highs()
    H.setOptionValue("output_flag", False)
    H.passModel(model)
    H.run()
    st = H.getModelStatus()
    name = str(st).split(".")[-1]
    if st != h.HighsModelStatus.kOptimal:
        return (False, None, name) if with_status else (False, None)
    z = np.array(H.getSolution().col_value)
    return (True, z, name) if with_status else (True, z)


def solve_highs(N, x0, xf, lim, dt, polys, sigma, force_final=True, with_status=False):
    """-> (feasible, cost, coeffs[N,12]).  cost = sum (6a)^2, i.e. Gurobi's ObjVal.  with_status=True appends HiGHS' model
    status name ("kOptimal", "kInfeasible", or whatever it stopped with: its QP solver sometimes gives up on these
    degenerate problems, which says nothing about feasibility)."""
    Q, Aeq, beq, Ain, bin_ = build(N, x0, xf, lim, dt, polys, sigma, force_final)
    ok, z, name = solve_qp_highs(Q, Aeq, beq, Ain, bin_, with_status=True)
    if not ok:
        return (False, np.nan, None, name) if with_status else (False, np.nan, None)
    cost = float(np.sum((6.0 * z.reshape(N, 12)[:, :3]) ** 2))
    return (True, cost, z.reshape(N, 12), name) if with_status else (True, cost, z.reshape(N, 12))
