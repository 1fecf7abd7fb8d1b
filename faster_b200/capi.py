"""ctypes binding of the C ABI in include/faster_b200.h (what a cgo/JNI-style binding of the reference side would
bind).  Used by tests and bench.py.  There is no CPU fallback: if the shared library is missing
or no GPU is present, calls raise.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FQ_LIB") or os.path.join(_HERE, "lib", "libfaster_b200.so")   # FQ_LIB: tuning variants
_lib = None

EXPORTS = ["fq_abi_version", "fq_create", "fq_destroy", "fq_last_error", "fq_set_option", "fq_solve_batch", "fq_solve_multi",
           "fq_solve_multi_async", "fq_wait", "fq_solve_multi_dev", "fq_gen_new_traj", "fq_gen_new_traj_sampled", "fq_gen_new_traj_exact", "fq_dt_initial", "fq_num_samples", "fq_fill_x",
           "fq_monotone_sigmas", "fq_plan_tables", "fq_ellipsoid_decomp", "fq_jps3d_plan", "fq_jps3d_plan_world", "fq_jps3d_rules"]


class FqError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FqError("%s is missing: run `python -m faster_b200.build` (no CPU fallback exists)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.fq_last_error.restype = C.c_char_p
        L.fq_last_error.argtypes = [C.c_void_p]
        L.fq_dt_initial.restype = C.c_double
        L.fq_dt_initial.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.fq_num_samples.restype = C.c_int
        L.fq_num_samples.argtypes = [C.c_int, C.c_double, C.c_double]
        L.fq_fill_x.restype = None
        L.fq_fill_x.argtypes = [C.c_int, C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_void_p]
        L.fq_monotone_sigmas.restype = C.c_long
        L.fq_monotone_sigmas.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_long]
        L.fq_plan_tables.restype = C.c_int
        L.fq_plan_tables.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.fq_ellipsoid_decomp.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_double, C.c_double,
                                          C.c_void_p, C.c_void_p, C.c_int]
        L.fq_create.argtypes = [C.POINTER(C.c_void_p), C.c_int]
        L.fq_destroy.argtypes = [C.c_void_p]
        L.fq_destroy.restype = None
        L.fq_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        L.fq_solve_batch.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 3 + [C.c_int, C.c_void_p,
                                     C.c_void_p, C.c_int] + [C.c_void_p] * 6
        L.fq_solve_multi.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 13
        L.fq_solve_multi_async.argtypes = L.fq_solve_multi.argtypes
        L.fq_wait.argtypes = [C.c_void_p]
        L.fq_solve_multi_dev.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 7 + \
                                        [C.c_int, C.c_int] + [C.c_void_p] * 7
        L.fq_gen_new_traj.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 3 + \
                                     [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int] + [C.c_void_p] * 5
        L.fq_gen_new_traj_sampled.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 3 + \
                                             [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                              C.c_double, C.c_int] + [C.c_void_p] * 6
        _lib = L
    return _lib


def _f64(x, n=None):
    x = np.ascontiguousarray(np.asarray(x, np.float64).reshape(-1))
    if n is not None and x.size != n:
        raise ValueError("expected %d doubles, got %d" % (n, x.size))
    return x


def pack_polys(polys):
    """list of (A[F,3], b[F]) -> (P, face_ofs int32[P+1], Ab float64[SF,4])."""
    ofs, rows = [0], []
    for A, b in polys:
        A = np.asarray(A, np.float64).reshape(-1, 3)
        b = np.asarray(b, np.float64).reshape(-1)
        rows.append(np.hstack([A, b[:, None]]))
        ofs.append(ofs[-1] + A.shape[0])
    Ab = np.ascontiguousarray(np.vstack(rows), np.float64) if rows else np.zeros((1, 4))
    return len(polys), np.asarray(ofs, np.int32), Ab


def dt_initial(x0, xf, lim, N):
    a, b, c = _f64(x0, 9), _f64(xf, 9), _f64(lim, 3)
    return lib().fq_dt_initial(a.ctypes.data, b.ctypes.data, c.ctypes.data, int(N))


def num_samples(N, dt, DC):
    return lib().fq_num_samples(int(N), float(dt), float(DC))


def fill_x(N, coeffs, dt, DC):
    n = num_samples(N, dt, DC)
    co = _f64(coeffs, 12 * N)
    out = np.zeros((n, 12))
    lib().fq_fill_x(int(N), co.ctypes.data, float(dt), float(DC), n, out.ctypes.data)
    return out


def monotone_sigmas(N, P):
    n = lib().fq_monotone_sigmas(int(N), int(P), None, 0)
    out = np.zeros((n, N), np.uint8)
    lib().fq_monotone_sigmas(int(N), int(P), out.ctypes.data, n)
    return out


def ellipsoid_decomp(path, obs, bbox=(2.0, 2.0, 1.0), inflate=0.42, z_ground=0.0, cap_rows=4096):
    """Host-side convex decomposition (fq_ellipsoid_decomp) -> list of (A[F,3], b[F]), one polytope per path segment."""
    path = np.ascontiguousarray(np.asarray(path, np.float64).reshape(-1, 3))
    obs = np.ascontiguousarray(np.asarray(obs, np.float64).reshape(-1, 3))
    n_seg = path.shape[0] - 1
    bb = _f64(bbox, 3)
    ofs = np.zeros(n_seg + 1, np.int32)
    Ab = np.zeros((cap_rows, 4))
    rows = lib().fq_ellipsoid_decomp(path.ctypes.data, n_seg, obs.ctypes.data if len(obs) else None, len(obs),
                                     bb.ctypes.data, float(inflate), float(z_ground), ofs.ctypes.data, Ab.ctypes.data,
                                     cap_rows)
    if rows < 0:
        raise FqError("fq_ellipsoid_decomp failed (%d)" % rows)
    return [(Ab[ofs[i]:ofs[i + 1], :3].copy(), Ab[ofs[i]:ofs[i + 1], 3].copy()) for i in range(n_seg)]


def jps3d_plan(grid, start, goal, use_jps=True, max_expand=-1):
    """grid: int8 array [zd, yd, xd] (x fastest in memory).  -> (path int[n,3] (x,y,z), cost in cells, expanded)."""
    g = np.ascontiguousarray(grid, np.int8)
    zd, yd, xd = g.shape
    s = np.asarray(start, np.int32).copy()
    t = np.asarray(goal, np.int32).copy()
    cap = 4 * (xd + yd + zd) + 16
    out = np.zeros((cap, 3), np.int32)
    cost, ex = C.c_double(np.inf), C.c_int(0)
    L = lib()
    L.fq_jps3d_plan.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                C.c_int, C.c_void_p, C.c_void_p]
    n = L.fq_jps3d_plan(g.ctypes.data, xd, yd, zd, s.ctypes.data, t.ctypes.data, int(use_jps), int(max_expand),
                        out.ctypes.data, cap, C.addressof(cost), C.addressof(ex))
    if n < 0:
        raise FqError("fq_jps3d_plan failed (%d)" % n)
    if n > cap:
        out = np.zeros((n, 3), np.int32)
        n = L.fq_jps3d_plan(g.ctypes.data, xd, yd, zd, s.ctypes.data, t.ctypes.data, int(use_jps), int(max_expand),
                            out.ctypes.data, n, C.addressof(cost), C.addressof(ex))
    return out[:n].copy(), cost.value, ex.value


def jps3d_plan_world(grid, origin, res, start, goal, use_jps=True, cap=4096):
    """World-coordinate plan with the reference's post-processing -> (path float[n,3], raw_cost in metres)."""
    g = np.ascontiguousarray(grid, np.int8)
    zd, yd, xd = g.shape
    o, s, t = _f64(origin, 3), _f64(start, 3), _f64(goal, 3)
    out = np.zeros((cap, 3))
    rc = C.c_double(np.inf)
    L = lib()
    L.fq_jps3d_plan_world.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p,
                                      C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    n = L.fq_jps3d_plan_world(g.ctypes.data, xd, yd, zd, o.ctypes.data, float(res), s.ctypes.data, t.ctypes.data,
                              int(use_jps), out.ctypes.data, cap, C.addressof(rc))
    if n < 0:
        raise FqError("fq_jps3d_plan_world failed (%d)" % n)
    return out[:n].copy(), rc.value


def jps3d_rules():
    ns = np.zeros((27, 3, 26), np.int32)
    f1 = np.zeros((27, 3, 12), np.int32)
    f2 = np.zeros((27, 3, 12), np.int32)
    cnt = np.zeros((27, 2), np.int32)
    L = lib()
    L.fq_jps3d_rules.argtypes = [C.c_void_p] * 4
    L.fq_jps3d_rules.restype = None
    L.fq_jps3d_rules(ns.ctypes.data, f1.ctypes.data, f2.ctypes.data, cnt.ctypes.data)
    return ns, f1, f2, cnt


def plan_tables(N, force_final):
    ne = 3 if force_final else 2
    nz, NY = N - ne, 6 * N + 1
    TZ = np.zeros((NY, max(nz, 0)))
    T0 = np.zeros((NY, 3 + ne))
    FT = np.zeros((ne, 3))
    r = lib().fq_plan_tables(int(N), int(bool(force_final)), TZ.ctypes.data, T0.ctypes.data, FT.ctypes.data)
    if r == 0:
        raise FqError("unsupported (N, force_final)")
    return TZ, T0, FT


class Solver:
    """Owns an fq_ctx on one GPU."""

    def __init__(self, device=0):
        self._L = lib()
        h = C.c_void_p()
        rc = self._L.fq_create(C.byref(h), int(device))
        if rc != 0:
            raise FqError("fq_create failed (%d): %s" % (rc, self._L.fq_last_error(None).decode()))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._L.fq_destroy(self._h)
            self._h = None

    __del__ = close

    def set_option(self, key, value):
        self._check(self._L.fq_set_option(self._h, key.encode(), int(value)))

    def _check(self, rc):
        if rc < 0:
            raise FqError("faster_b200 error %d: %s" % (rc, self._L.fq_last_error(self._h).decode()))
        return rc

    def solve_batch(self, N, x0, xf, lim, polys, dts, sigmas, force_final=True, want_coeffs=False, want_iters=False):
        P, ofs, Ab = pack_polys(polys)
        dts = _f64(dts)
        n = dts.size
        sig = np.ascontiguousarray(np.asarray(sigmas, np.uint8).reshape(n, N)) if P > 0 else np.zeros((n, N), np.uint8)
        x0, xf, lim = _f64(x0, 9), _f64(xf, 9), _f64(lim, 3)
        feas = np.zeros(n, np.uint8)
        cost = np.zeros(n)
        co = np.zeros((n, N, 12)) if want_coeffs else None
        it = np.zeros(n, np.int32) if want_iters else None
        self._check(self._L.fq_solve_batch(self._h, int(N), int(bool(force_final)), x0.ctypes.data, xf.ctypes.data,
                                           lim.ctypes.data, P, ofs.ctypes.data, Ab.ctypes.data, n, dts.ctypes.data,
                                           sig.ctypes.data, feas.ctypes.data, cost.ctypes.data,
                                           co.ctypes.data if want_coeffs else None,
                                           it.ctypes.data if want_iters else None))
        return feas, cost, co, it

    def solve_multi(self, N, force_final, x0, xf, lim, poly_ofs, face_ofs, Ab, cand_ofs, dts, sigmas,
                    want_coeffs=False, want_iters=False, out=None, deferred=False):
        """Arrays as in fq_solve_multi; numpy (possibly pinned torch-backed) host arrays.  deferred=True uses
        fq_solve_multi_async: the outputs are valid after wait() (keep every array alive until then)."""
        n_prob = len(cand_ofs) - 1
        n = int(cand_ofs[-1])
        if out is None:
            feas = np.zeros(n, np.uint8)
            cost = np.zeros(n)
            co = np.zeros((n, N, 12)) if want_coeffs else None
            it = np.zeros(n, np.int32) if want_iters else None
        else:
            feas, cost, co, it = out
        fn = self._L.fq_solve_multi_async if deferred else self._L.fq_solve_multi
        self._check(fn(self._h, int(N), int(bool(force_final)), n_prob, x0.ctypes.data,
                       xf.ctypes.data, lim.ctypes.data, poly_ofs.ctypes.data, face_ofs.ctypes.data,
                       Ab.ctypes.data, cand_ofs.ctypes.data, dts.ctypes.data, sigmas.ctypes.data,
                       feas.ctypes.data, cost.ctypes.data,
                       co.ctypes.data if co is not None else None,
                       it.ctypes.data if it is not None else None))
        return feas, cost, co, it

    def wait(self):
        """fq_wait: drain everything enqueued on the context's own streams (deferred batches included)."""
        self._check(self._L.fq_wait(self._h))

    def solve_multi_dev(self, N, force_final, n_prob, d_x0, d_xf, d_lim, d_poly_ofs, d_face_ofs, d_Ab, d_cand_ofs,
                        max_cand, max_faces, d_dt, d_sigma, d_feas, d_cost, d_coeffs=0, d_iters=0, stream=0):
        """All d_* are integer device addresses (e.g. torch tensor .data_ptr())."""
        self._check(self._L.fq_solve_multi_dev(self._h, int(N), int(bool(force_final)), int(n_prob), d_x0, d_xf,
                                               d_lim, d_poly_ofs, d_face_ofs, d_Ab, d_cand_ofs, int(max_cand),
                                               int(max_faces), d_dt, d_sigma, d_feas, d_cost, d_coeffs or None,
                                               d_iters or None, stream or None))

    def gen_new_traj_sampled(self, N, x0, xf, lim, polys, dts, sigmas, DC, force_final=True, max_samples=8192):
        """gen_new_traj + fillX on the device -> dict(solved, dt_index, sigma_index, cost, coeffs, X[n,12])."""
        P, ofs, Ab = pack_polys(polys)
        dts = _f64(dts)
        sig = np.ascontiguousarray(np.asarray(sigmas, np.uint8).reshape(-1, N)) if P > 0 else np.zeros((1, N), np.uint8)
        x0, xf, lim = _f64(x0, 9), _f64(xf, 9), _f64(lim, 3)
        di, si, cost, ns = C.c_int(-1), C.c_int(-1), C.c_double(np.inf), C.c_int(0)
        co = np.zeros((N, 12))
        X = np.zeros((max_samples, 12))
        rc = self._check(self._L.fq_gen_new_traj_sampled(self._h, int(N), int(bool(force_final)), x0.ctypes.data,
                                                         xf.ctypes.data, lim.ctypes.data, P, ofs.ctypes.data,
                                                         Ab.ctypes.data, dts.size, dts.ctypes.data, sig.shape[0],
                                                         sig.ctypes.data, float(DC), int(max_samples), C.addressof(di),
                                                         C.addressof(si), C.addressof(cost), co.ctypes.data, X.ctypes.data,
                                                         C.addressof(ns)))
        return dict(solved=bool(rc), dt_index=di.value, sigma_index=si.value, cost=cost.value, coeffs=co, X=X[:ns.value])

    def gen_new_traj_exact(self, N, x0, xf, lim, polys, dts, force_final=True):
        """Exact MIQP sweep (branch-and-bound over all assignments) -> dict(solved, dt_index, sigma, cost, coeffs, nodes, exact)."""
        P, ofs, Ab = pack_polys(polys)
        dts = _f64(dts)
        x0, xf, lim = _f64(x0, 9), _f64(xf, 9), _f64(lim, 3)
        di, cost, nodes, exact = C.c_int(-1), C.c_double(np.inf), C.c_long(0), C.c_int(0)
        sig = np.zeros(N, np.uint8)
        co = np.zeros((N, 12))
        self._L.fq_gen_new_traj_exact.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 3 + [C.c_int, C.c_void_p, C.c_void_p,
                                                   C.c_int] + [C.c_void_p] * 7
        rc = self._check(self._L.fq_gen_new_traj_exact(self._h, int(N), int(bool(force_final)), x0.ctypes.data, xf.ctypes.data,
                                                       lim.ctypes.data, P, ofs.ctypes.data, Ab.ctypes.data, dts.size,
                                                       dts.ctypes.data, C.addressof(di), sig.ctypes.data, C.addressof(cost),
                                                       co.ctypes.data, C.addressof(nodes), C.addressof(exact)))
        return dict(solved=bool(rc), dt_index=di.value, sigma=sig, cost=cost.value, coeffs=co, nodes=nodes.value,
                    exact=bool(exact.value))

    def gen_new_traj(self, N, x0, xf, lim, polys, dts, sigmas, force_final=True):
        """-> dict(solved, dt_index, sigma_index, cost, coeffs[N,12])."""
        P, ofs, Ab = pack_polys(polys)
        dts = _f64(dts)
        sig = np.ascontiguousarray(np.asarray(sigmas, np.uint8).reshape(-1, N)) if P > 0 else np.zeros((1, N), np.uint8)
        x0, xf, lim = _f64(x0, 9), _f64(xf, 9), _f64(lim, 3)
        di, si, cost = C.c_int(-1), C.c_int(-1), C.c_double(np.inf)
        co = np.zeros((N, 12))
        rc = self._check(self._L.fq_gen_new_traj(self._h, int(N), int(bool(force_final)), x0.ctypes.data,
                                                 xf.ctypes.data, lim.ctypes.data, P, ofs.ctypes.data, Ab.ctypes.data,
                                                 dts.size, dts.ctypes.data, sig.shape[0], sig.ctypes.data,
                                                 C.addressof(di), C.addressof(si), C.addressof(cost), co.ctypes.data))
        return dict(solved=bool(rc), dt_index=di.value, sigma_index=si.value, cost=cost.value, coeffs=co)
