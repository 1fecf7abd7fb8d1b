"""numpy mirror of the CUDA kernel's algorithm (faster_b200/csrc/fq_kernels.cu), step for step, on the same plan
tables.  Test infrastructure: lets the CPU suite check the kernel's mathematics (normalised variables, table rows,
Householder/Givens updates, ratio test) against the oracle without a GPU.  Not used by the product.
"""
import numpy as np

TOL = 1e-8
EPS_DEP = 1e-18
ZZ_FLOOR = 1e-30
MAX_ITERS = 400


def solve(tables, N, x0, xf, lim, dt, polys, sigma, force_final=True, normalised=False, thin=False, trace=None):
    """-> (status, cost, coeffs[N,12], iters).  normalised=False mirrors the size-generic kernel (entering row = largest
    violation); normalised=True mirrors the size-specialised kernel (largest (violation - tol) / |TZ[y]|).
    thin=True: the thin factorisation of the specialised kernel (J1 = orthonormal basis of the active normals only,
    Gram-Schmidt with one re-orthogonalisation pass when the residual is small) instead of the full orthogonal J.
    trace: a list that receives, at every scan for the entering row, (w, length of the path w travelled since the previous
    scan) -- tools/scan_skip_model.py studies how much of a scan a distance bound could skip."""
    TZ, T0, FT = tables
    ne = 3 if force_final else 2
    nz, NY = N - ne, 6 * N + 1
    nw = 3 * nz
    x0 = np.asarray(x0, float)
    xf = np.asarray(xf, float)
    inv = np.array([1 / dt, 1 / dt ** 2, 1 / dt ** 3])
    Yeq = np.zeros((3, NY))
    for ax in range(3):
        s0 = np.array([x0[ax], x0[3 + ax] * dt, x0[6 + ax] * dt * dt])
        tgt = ([xf[ax]] if force_final else []) + [xf[3 + ax] * dt, xf[6 + ax] * dt * dt]
        rhs = np.array(tgt) - FT @ s0
        Yeq[ax] = T0[:, :3] @ s0 + T0[:, 3:] @ rhs
    w = np.zeros(nw)
    J = np.zeros((nw, nw)) if thin else np.eye(nw)
    R = np.zeros((nw, nw))
    lam = np.zeros(nw)
    q = 0
    it = 0

    def Yof(w):
        return Yeq + (TZ @ w.reshape(3, nz).T).T if nz > 0 else Yeq.copy()

    Y = Yof(w)
    rows = []                                    # (t, A row, b) per corridor row
    if len(polys):
        for t in range(N):
            A, b = polys[int(sigma[t])]
            for f in range(len(b)):
                rows.append((t, np.asarray(A[f], float), float(b[f])))
    TZN = np.sum(TZ * TZ, axis=1) if nz > 0 else np.zeros(NY)
    SY = np.where(TZN > 1e-30, 1.0 / np.sqrt(np.maximum(TZN, 1e-300)), 1e15)

    def rank(viol, y, scale=1.0):
        return (viol - TOL) * SY[y] / scale if normalised else viol - TOL

    travelled = 0.0
    while True:
        if trace is not None:
            trace.append((w.copy(), travelled))
            travelled = 0.0
        best, desc = 0.0, None
        for typ in range(3):
            for ax in range(3):
                for t in range(N):
                    y = (typ + 1) * N + 1 + t
                    val = Y[ax, y]
                    viol = abs(val) * inv[typ] - lim[typ]
                    if rank(viol, y, inv[typ]) > best:
                        wv = np.zeros(3)
                        wv[ax] = inv[typ] if val > 0 else -inv[typ]
                        best, desc = rank(viol, y, inv[typ]), (y, wv, lim[typ])
        for t, a, b in rows:
            for y in (t, 4 * N + 1 + t, 5 * N + 1 + t, t + 1):
                v = a @ Y[:, y] - b
                if rank(v, y) > best:
                    best, desc = rank(v, y), (y, a, b)
        if desc is None:
            status = 1
            break
        y, wv, h = desc
        g = np.concatenate([wv[ax] * TZ[y] for ax in range(3)]) if nz > 0 else np.zeros(0)
        gg = g @ g
        lam_p = 0.0
        done = False
        while True:
            it += 1
            if it > MAX_ITERS:
                status, done = -1, True
                break
            viol = wv @ Y[:, y] - h
            d = J.T @ g
            if thin:
                d[q:] = 0.0
                z = -(g - J[:, :q] @ d[:q])
                zz = z @ z
                if zz < 0.01 * gg and q > 0:                       # re-orthogonalise once ("twice is enough")
                    c = J[:, :q].T @ z
                    z = z - J[:, :q] @ c
                    d[:q] = d[:q] - c                              # g = J1 (d1 - c) + (-z): keep the R column consistent
                    zz = z @ z
            else:
                zz = d[q:] @ d[q:]
                z = -J[:, q:] @ d[q:]
            r = np.linalg.solve(np.triu(R[:q, :q]), d[:q]) if q > 0 else np.zeros(0)
            dep = zz <= max(EPS_DEP * gg, ZZ_FLOOR)
            t1, l = np.inf, -1
            for k in range(q):
                if r[k] > 0 and lam[k] / r[k] < t1:
                    t1, l = lam[k] / r[k], k
            t2 = np.inf if dep else viol / zz
            if t1 == np.inf and t2 == np.inf:
                status, done = 0, True
                break
            if t2 <= t1:
                w = w + t2 * z
                travelled += abs(t2) * np.sqrt(zz)
                lam[:q] -= t2 * r
                lam_p += t2
                nrm = np.sqrt(zz)
                if thin:
                    J[:, q] = -z / nrm                            # new basis vector: the normalised residual of g
                    R[:q, q] = d[:q]
                    R[q, q] = nrm
                else:
                    dq = d[q]
                    sgn = 1.0 if dq >= 0 else -1.0
                    beta = 1.0 / (zz + abs(dq) * nrm)
                    v = d[q:].copy()
                    v[0] += sgn * nrm
                    u = -z + sgn * nrm * J[:, q]
                    J[:, q:] -= beta * np.outer(u, v)
                    R[:q, q] = d[:q]
                    R[q, q] = -sgn * nrm
                lam[q] = lam_p
                q += 1
                Y = Yof(w)
                break
            if not dep:
                w = w + t1 * z
                travelled += abs(t1) * np.sqrt(zz)
            lam[:q] -= t1 * r
            lam_p += t1
            # drop l
            R[:q, l:q - 1] = R[:q, l + 1:q]
            R[:, q - 1] = 0
            lam[l:q - 1] = lam[l + 1:q]
            for j in range(l, q - 1):
                p, s = R[j, j], R[j + 1, j]
                hh = np.hypot(p, s)
                c, sn = (p / hh, s / hh) if hh > 0 else (1.0, 0.0)
                Rj, Rj1 = R[j, j:q - 1].copy(), R[j + 1, j:q - 1].copy()
                R[j, j:q - 1] = c * Rj + sn * Rj1
                R[j + 1, j:q - 1] = c * Rj1 - sn * Rj
                Jj, Jj1 = J[:, j].copy(), J[:, j + 1].copy()
                J[:, j] = c * Jj + sn * Jj1
                J[:, j + 1] = c * Jj1 - sn * Jj
            q -= 1
            if not dep:
                Y = Yof(w)
        if done:
            break
    if status != 1:
        return status, np.inf, None, it
    U = Y[:, 3 * N + 1:4 * N + 1]
    cost = float(np.sum(U * U) * inv[2] ** 2)
    co = np.zeros((N, 12))
    for t in range(N):
        for ax in range(3):
            co[t, ax] = Y[ax, 3 * N + 1 + t] * inv[2] / 6.0
            co[t, 3 + ax] = Y[ax, 2 * N + 1 + t] * inv[1] / 2.0
            co[t, 6 + ax] = Y[ax, N + 1 + t] * inv[0]
            co[t, 9 + ax] = Y[ax, t]
    return status, cost, co, it
