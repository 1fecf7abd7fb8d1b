"""TEST INFRASTRUCTURE ONLY -- the REFERENCE's own convex decomposition (DecompUtil's EllipsoidDecomp3D driven like
JPS_Manager::cvxEllipsoidDecomp, faster/src/jps_manager.cpp:80-127), compiled from /root/reference by oracle/Makefile into
oracle/_ref/libdecomp_ref.so (oracle/decomp_ref_wrap.cpp; oracle/stub_eigen stands in for Eigen).  Only tests/ import it."""
import ctypes as C
import os

import numpy as np

_SO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libdecomp_ref.so")
_lib = None


def available():
    return os.path.exists(_SO)


def cvx_ellipsoid_decomp(path, obs, bbox=(2.0, 2.0, 1.0), inflate=0.42, z_ground=0.0, cap_rows=8192):
    """-> list of (A[F,3], b[F]), one polytope per path segment, rows in the reference's order (obstacle faces, the six
    local-bbox faces, the ground face)."""
    global _lib
    if _lib is None:
        _lib = C.CDLL(_SO)
        _lib.decompref_cvx.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_double, C.c_double, C.c_void_p,
                                       C.c_void_p, C.c_int]
        _lib.decompref_cvx.restype = C.c_int
    path = np.ascontiguousarray(np.asarray(path, np.float64).reshape(-1, 3))
    obs = np.ascontiguousarray(np.asarray(obs, np.float64).reshape(-1, 3))
    n = path.shape[0]
    bb = np.ascontiguousarray(np.asarray(bbox, np.float64))
    ofs = np.zeros(n, np.int32)
    Ab = np.zeros((cap_rows, 4))
    rows = _lib.decompref_cvx(path.ctypes.data, n, obs.ctypes.data if len(obs) else None, len(obs), bb.ctypes.data, float(inflate),
                              float(z_ground), ofs.ctypes.data, Ab.ctypes.data, cap_rows)
    if rows < 0:
        raise RuntimeError("decompref_cvx: cap_rows too small")
    return [(Ab[ofs[i]:ofs[i + 1], :3].copy(), Ab[ofs[i]:ofs[i + 1], 3].copy()) for i in range(n - 1)]
