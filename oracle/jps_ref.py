"""TEST INFRASTRUCTURE ONLY -- loader for oracle/_ref/libjps_ref.so: the REFERENCE's own JPS3D graph search
(thirdparty/jps3d/src/jps_planner/graph_search.cpp) compiled from /root/reference by oracle/Makefile with a stub for
boost::heap (oracle/stub_boost).  Exists only where it was built (this container; it travels to the GPU box prebuilt)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libjps_ref.so")


def available():
    if not os.path.exists(_SO) and os.path.exists("/root/reference/thirdparty/jps3d/src/jps_planner/graph_search.cpp"):
        subprocess.call(["make", "-C", _HERE, "-s"])
    return os.path.exists(_SO)


def plan(grid, start, goal, use_jps=True, max_expand=-1):
    """grid int8 [zd,yd,xd] -> (path int[n,3], cost in cells, closed-set size)."""
    L = C.CDLL(_SO)
    g = np.ascontiguousarray(grid, np.int8)
    zd, yd, xd = g.shape
    cap = xd * yd + 64
    out = np.zeros((cap, 3), np.int32)
    cost, nc = C.c_double(np.inf), C.c_int(0)
    n = L.jpsref_plan(g.ctypes.data_as(C.c_char_p), xd, yd, zd, int(start[0]), int(start[1]), int(start[2]), int(goal[0]),
                      int(goal[1]), int(goal[2]), int(use_jps), int(max_expand), out.ctypes.data_as(C.c_void_p), cap,
                      C.byref(cost), C.byref(nc))
    return out[:min(n, cap)].copy(), cost.value, nc.value


def tables():
    L = C.CDLL(_SO)
    ns = np.zeros((27, 3, 26), np.int32)
    f1 = np.zeros((27, 3, 12), np.int32)
    f2 = np.zeros((27, 3, 12), np.int32)
    L.jpsref_tables(ns.ctypes.data_as(C.c_void_p), f1.ctypes.data_as(C.c_void_p), f2.ctypes.data_as(C.c_void_p))
    return ns, f1, f2
