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
           "fq_monotone_sigmas", "fq_plan_tables", "fq_ellipsoid_decomp", "fq_jps3d_plan", "fq_jps3d_plan_world", "fq_jps3d_rules",
           "fq_replan_pairs", "fq_replan_pairs_async", "fq_replan_pairs_dev", "fq_create_multi", "fq_comm_unique_id", "fq_comm_init",
           "fq_comm_info", "fq_allgather_dev", "fq_shard_range", "fq_solve_multi_sharded", "fq_solve_batch_cert", "fq_has_feature"]


class FqError(RuntimeError):
    pass


class PairResult(C.Structure):
    """fq_pair_result of include/faster_b200.h (144 bytes)."""
    _fields_ = [("whole_dt_index", C.c_int), ("whole_sigma_index", C.c_int), ("safe_dt_index", C.c_int),
                ("safe_sigma_index", C.c_int), ("whole_cost", C.c_double), ("safe_cost", C.c_double),
                ("whole_dt", C.c_double), ("safe_dt", C.c_double), ("whole_dt_base", C.c_double),
                ("safe_dt_base", C.c_double), ("n_samples_whole", C.c_int), ("k_safe", C.c_int), ("R", C.c_double * 9)]


PAIR_RESULT_DTYPE = np.dtype([("whole_dt_index", np.int32), ("whole_sigma_index", np.int32), ("safe_dt_index", np.int32),
                              ("safe_sigma_index", np.int32), ("whole_cost", np.float64), ("safe_cost", np.float64),
                              ("whole_dt", np.float64), ("safe_dt", np.float64), ("whole_dt_base", np.float64),
                              ("safe_dt_base", np.float64), ("n_samples_whole", np.int32), ("k_safe", np.int32),
                              ("R", np.float64, (9,))])
assert PAIR_RESULT_DTYPE.itemsize == C.sizeof(PairResult) == 144


class PairArgs(C.Structure):
    """fq_pair_args of include/faster_b200.h; pointers as integers (host or device addresses)."""
    _fields_ = [("n_prob", C.c_int), ("N_whole", C.c_int), ("N_safe", C.c_int), ("DC", C.c_double), ("r_fraction", C.c_double),
                ("x0", C.c_void_p), ("xf_whole", C.c_void_p), ("xf_safe", C.c_void_p), ("lim", C.c_void_p),
                ("poly_ofs_whole", C.c_void_p), ("face_ofs_whole", C.c_void_p), ("Ab_whole", C.c_void_p),
                ("poly_ofs_safe", C.c_void_p), ("face_ofs_safe", C.c_void_p), ("Ab_safe", C.c_void_p),
                ("n_fac_whole", C.c_int), ("factors_whole", C.c_void_p), ("n_sig_whole", C.c_int), ("sigmas_whole", C.c_void_p),
                ("n_fac_safe", C.c_int), ("factors_safe", C.c_void_p), ("n_sig_safe", C.c_int), ("sigmas_safe", C.c_void_p),
                ("feasible_whole", C.c_void_p), ("cost_whole", C.c_void_p), ("feasible_safe", C.c_void_p), ("cost_safe", C.c_void_p),
                ("coeffs_whole", C.c_void_p), ("coeffs_safe", C.c_void_p), ("results", C.c_void_p),
                ("max_faces_whole", C.c_int), ("max_poly_faces_whole", C.c_int), ("max_faces_safe", C.c_int),
                ("max_poly_faces_safe", C.c_int)]


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
        L.fq_replan_pairs.argtypes = [C.c_void_p, C.POINTER(PairArgs)]
        L.fq_replan_pairs_async.argtypes = [C.c_void_p, C.POINTER(PairArgs)]
        L.fq_replan_pairs_dev.argtypes = [C.c_void_p, C.POINTER(PairArgs), C.c_void_p, C.c_void_p]
        L.fq_create_multi.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_void_p]
        L.fq_comm_unique_id.argtypes = [C.c_void_p]
        L.fq_comm_init.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.fq_comm_info.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.fq_allgather_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_void_p]
        L.fq_shard_range.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.fq_solve_multi_sharded.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 13
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


def has_feature(name):
    L = lib()
    L.fq_has_feature.argtypes = [C.c_char_p]
    return bool(L.fq_has_feature(name.encode()))


def shard_range(n_prob, cand_ofs, rank, world):
    """fq_shard_range -> (lo, hi): the problems of `rank`."""
    lo, hi = C.c_int(0), C.c_int(0)
    co = None if cand_ofs is None else np.ascontiguousarray(cand_ofs, np.int32)
    rc = lib().fq_shard_range(int(n_prob), co.ctypes.data if co is not None else None, int(rank), int(world),
                              C.addressof(lo), C.addressof(hi))
    if rc != 0:
        raise FqError("fq_shard_range: bad arguments")
    return lo.value, hi.value


def comm_unique_id():
    """128-byte NCCL id (rank 0 creates it, the launcher distributes it)."""
    buf = (C.c_char * 128)()
    rc = lib().fq_comm_unique_id(C.addressof(buf))
    if rc != 0:
        raise FqError("fq_comm_unique_id failed (%d): %s" % (rc, lib().fq_last_error(None).decode()))
    return bytes(buf)


def make_pair_workload(whole, safe, factors_whole, sigmas_whole, factors_safe, sigmas_safe, DC=0.01, r_fraction=0.6):
    """Packs two lists of corridor dicts (corridor.make_* output; whole[j] and safe[j] belong to one replan) into the
    host arrays of fq_pair_args.  -> dict of numpy arrays + sizes."""
    n = len(whole)
    assert len(safe) == n

    def csr(probs):
        po, fo, rows = [0], [0], []
        for pb in probs:
            for A, b in pb["polys"]:
                rows.append(np.hstack([np.asarray(A, float).reshape(-1, 3), np.asarray(b, float).reshape(-1, 1)]))
                fo.append(fo[-1] + len(b))
            po.append(po[-1] + len(pb["polys"]))
        Ab = np.ascontiguousarray(np.vstack(rows)) if rows else np.zeros((1, 4))
        fo = np.asarray(fo, np.int32)
        po = np.asarray(po, np.int32)
        mf = int(max(fo[po[j + 1]] - fo[po[j]] for j in range(len(probs)))) if rows else 1
        mpf = int(np.diff(fo).max()) if len(fo) > 1 else 0
        return po, fo, Ab, mf, mpf
    pw, fw, Aw, mfw, mpfw = csr(whole)
    ps, fs, As, mfs, mpfs = csr(safe)
    Nw, Ns = whole[0]["N"], safe[0]["N"]
    sw = np.ascontiguousarray(np.asarray(sigmas_whole, np.uint8).reshape(-1, Nw))
    ss = np.ascontiguousarray(np.asarray(sigmas_safe, np.uint8).reshape(-1, Ns))
    return dict(n_prob=n, N_whole=Nw, N_safe=Ns, DC=float(DC), r_fraction=float(r_fraction),
                x0=np.ascontiguousarray([pb["x0"] for pb in whole], np.float64),
                xf_whole=np.ascontiguousarray([pb["xf"] for pb in whole], np.float64),
                xf_safe=np.ascontiguousarray([pb["xf"] for pb in safe], np.float64),
                lim=np.ascontiguousarray([pb["lim"] for pb in whole], np.float64),
                poly_ofs_whole=pw, face_ofs_whole=fw, Ab_whole=Aw, poly_ofs_safe=ps, face_ofs_safe=fs, Ab_safe=As,
                factors_whole=_f64(factors_whole), sigmas_whole=sw, factors_safe=_f64(factors_safe), sigmas_safe=ss,
                max_faces_whole=mfw, max_poly_faces_whole=mpfw, max_faces_safe=mfs, max_poly_faces_safe=mpfs)


PAIR_INPUT_KEYS = ["x0", "xf_whole", "xf_safe", "lim", "poly_ofs_whole", "face_ofs_whole", "Ab_whole", "poly_ofs_safe",
                   "face_ofs_safe", "Ab_safe", "factors_whole", "sigmas_whole", "factors_safe", "sigmas_safe"]


def pair_args(w, ptr, out_ptrs):
    """PairArgs from a make_pair_workload dict; ptr(key) -> address of input array `key`, out_ptrs: dict of output
    addresses (feasible_whole, cost_whole, feasible_safe, cost_safe, coeffs_whole, coeffs_safe, results; 0/None = NULL)."""
    a = PairArgs()
    a.n_prob, a.N_whole, a.N_safe, a.DC, a.r_fraction = w["n_prob"], w["N_whole"], w["N_safe"], w["DC"], w["r_fraction"]
    for k in PAIR_INPUT_KEYS:
        setattr(a, k, ptr(k))
    a.n_fac_whole, a.n_sig_whole = len(w["factors_whole"]), len(w["sigmas_whole"])
    a.n_fac_safe, a.n_sig_safe = len(w["factors_safe"]), len(w["sigmas_safe"])
    for k in ("feasible_whole", "cost_whole", "feasible_safe", "cost_safe", "coeffs_whole", "coeffs_safe", "results"):
        setattr(a, k, out_ptrs.get(k) or None)
    for k in ("max_faces_whole", "max_poly_faces_whole", "max_faces_safe", "max_poly_faces_safe"):
        setattr(a, k, int(w[k]))
    return a


class Solver:
    """Owns an fq_ctx on one GPU (or, with n_gpus > 1, a single-process multi-GPU group: fq_create_multi)."""

    def __init__(self, device=0, n_gpus=1, devices=None):
        self._L = lib()
        h = C.c_void_p()
        if n_gpus > 1 or devices is not None:
            dv = None if devices is None else np.ascontiguousarray(devices, np.int32)
            n = n_gpus if devices is None else len(dv)
            rc = self._L.fq_create_multi(C.byref(h), int(n), dv.ctypes.data if dv is not None else None)
        else:
            rc = self._L.fq_create(C.byref(h), int(device))
        if rc != 0:
            raise FqError("fq_create failed (%d): %s" % (rc, self._L.fq_last_error(None).decode()))
        self._h = h

    def comm_init(self, id128, rank, world):
        """Attach this (single-GPU) context to a communicator: one process per GPU."""
        buf = C.create_string_buffer(bytes(id128), 128)
        self._check(self._L.fq_comm_init(self._h, C.addressof(buf), int(rank), int(world)))

    def comm_info(self):
        r, w, v = C.c_int(0), C.c_int(1), C.c_int(0)
        self._check(self._L.fq_comm_info(self._h, C.addressof(r), C.addressof(w), C.addressof(v)))
        return r.value, w.value, v.value

    def allgather_dev(self, d_send, d_recv, nbytes, stream=0):
        self._check(self._L.fq_allgather_dev(self._h, d_send, d_recv, int(nbytes), stream or None))

    def replan_pairs(self, w, want_candidates=True, want_coeffs=True, deferred=False, out=None):
        """fq_replan_pairs on a make_pair_workload dict (host arrays).  -> dict(results (structured array), feasible_whole,
        cost_whole, feasible_safe, cost_safe, coeffs_whole, coeffs_safe).  deferred: valid after wait(); `out` reuses a
        previous return value's arrays (keep it and `w` alive until then)."""
        n = w["n_prob"]
        ncw = n * len(w["factors_whole"]) * len(w["sigmas_whole"])
        ncs = n * len(w["factors_safe"]) * len(w["sigmas_safe"])
        if out is None:
            out = dict(results=np.zeros(n, PAIR_RESULT_DTYPE))
            if want_candidates:
                out.update(feasible_whole=np.zeros(ncw, np.uint8), cost_whole=np.zeros(ncw), feasible_safe=np.zeros(ncs, np.uint8),
                           cost_safe=np.zeros(ncs))
            if want_coeffs:
                out.update(coeffs_whole=np.zeros((n, w["N_whole"], 12)), coeffs_safe=np.zeros((n, w["N_safe"], 12)))
        a = pair_args(w, lambda k: w[k].ctypes.data, {k: v.ctypes.data for k, v in out.items()})
        fn = self._L.fq_replan_pairs_async if deferred else self._L.fq_replan_pairs
        self._check(fn(self._h, C.byref(a)))
        return out

    def replan_pairs_dev(self, a, results_all=0, stream=0):
        """fq_replan_pairs_dev: `a` is a PairArgs holding device addresses."""
        self._check(self._L.fq_replan_pairs_dev(self._h, C.byref(a), results_all or None, stream or None))

    def solve_multi_sharded(self, N, force_final, x0, xf, lim, poly_ofs, face_ofs, Ab, cand_ofs, dts, sigmas):
        """-> (feasible, cost, win_idx, win_cost); see fq_solve_multi_sharded."""
        n_prob = len(cand_ofs) - 1
        n = int(cand_ofs[-1])
        feas = np.zeros(n, np.uint8)
        cost = np.full(n, np.nan)
        wi = np.zeros(n_prob, np.int32)
        wc = np.zeros(n_prob)
        self._check(self._L.fq_solve_multi_sharded(self._h, int(N), int(bool(force_final)), n_prob, x0.ctypes.data, xf.ctypes.data,
                                                   lim.ctypes.data, poly_ofs.ctypes.data, face_ofs.ctypes.data, Ab.ctypes.data,
                                                   cand_ofs.ctypes.data, dts.ctypes.data, sigmas.ctypes.data, feas.ctypes.data,
                                                   cost.ctypes.data, wi.ctypes.data, wc.ctypes.data))
        return feas, cost, wi, wc

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

    def solve_batch_cert(self, N, x0, xf, lim, polys, dts, sigmas, force_final=True):
        """fq_solve_batch_cert -> (feasible, cost, cert[n, stride]): Farkas certificates of the infeasible candidates."""
        P, ofs, Ab = pack_polys(polys)
        dts = _f64(dts)
        n = dts.size
        sig = np.ascontiguousarray(np.asarray(sigmas, np.uint8).reshape(n, N)) if P > 0 else np.zeros((n, N), np.uint8)
        x0, xf, lim = _f64(x0, 9), _f64(xf, 9), _f64(lim, 3)
        feas = np.zeros(n, np.uint8)
        cost = np.zeros(n)
        stride = 4 + 6 * 16
        cert = np.zeros((n, stride))
        self._L.fq_solve_batch_cert.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 3 + [C.c_int, C.c_void_p, C.c_void_p,
                                                C.c_int] + [C.c_void_p] * 5 + [C.c_int]
        self._check(self._L.fq_solve_batch_cert(self._h, int(N), int(bool(force_final)), x0.ctypes.data, xf.ctypes.data, lim.ctypes.data,
                                                P, ofs.ctypes.data, Ab.ctypes.data, n, dts.ctypes.data, sig.ctypes.data,
                                                feas.ctypes.data, cost.ctypes.data, cert.ctypes.data, stride))
        return feas, cost, cert

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
