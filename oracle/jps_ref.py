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


# ---- the planner layer above the graph search: JPSPlanner<3>::plan over JPS::MapUtil<3> (jps_planner.cpp, map_util.h), compiled
#      unmodified from /root/reference into oracle/_ref/libjpsplan_ref.so (oracle/jpsplan_ref_wrap.cpp; stub_eigen / stub_ros /
#      stub_pcl stand in for Eigen, ROS and PCL)
_PLAN_SO = os.path.join(_HERE, "_ref", "libjpsplan_ref.so")


def planner_available():
    if not os.path.exists(_PLAN_SO) and os.path.exists("/root/reference/thirdparty/jps3d/src/jps_planner/jps_planner.cpp"):
        subprocess.call(["make", "-C", _HERE, "-s"])
    return os.path.exists(_PLAN_SO)


def plan_world(grid, origin, res, start, goal, use_jps=True, cap=8192):
    """The reference's world-coordinate plan with its path simplification.  grid int8 [zd,yd,xd] (0 free, 100 occupied,
    -1 unknown) -> (path float[n,3] start -> goal, raw path float[m,3], planner status: 0 ok, 1 start not free, 2 goal not
    free, -1 no path)."""
    L = C.CDLL(_PLAN_SO)
    L.jpsplanref_plan.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_int,
                                  C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    g = np.ascontiguousarray(grid, np.int8)
    zd, yd, xd = g.shape
    o, s, t = (np.ascontiguousarray(np.asarray(v, np.float64)) for v in (origin, start, goal))
    out, raw = np.zeros((cap, 3)), np.zeros((cap, 3))
    nr, st = C.c_int(0), C.c_int(0)
    n = L.jpsplanref_plan(g.ctypes.data, xd, yd, zd, o.ctypes.data, float(res), s.ctypes.data, t.ctypes.data, int(use_jps),
                          out.ctypes.data, cap, raw.ctypes.data, C.addressof(nr), C.addressof(st))
    return out[:min(n, cap)].copy(), raw[:min(nr.value, cap)].copy(), st.value


def blocked(grid, origin, res, p1, p2):
    """MapUtil::isBlocked (map_util.h:371-383): a ray-traced cell between p1 and p2 is occupied."""
    L = C.CDLL(_PLAN_SO)
    L.jpsplanref_blocked.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
    g = np.ascontiguousarray(grid, np.int8)
    zd, yd, xd = g.shape
    o, a, b = (np.ascontiguousarray(np.asarray(v, np.float64)) for v in (origin, p1, p2))
    return bool(L.jpsplanref_blocked(g.ctypes.data, xd, yd, zd, o.ctypes.data, float(res), a.ctypes.data, b.ctypes.data))
