"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the convex decomposition FASTER runs before the solver:
JPS_Manager::cvxEllipsoidDecomp (faster/src/jps_manager.cpp:80-127) over DecompUtil's EllipsoidDecomp3D
(thirdparty/DecompROS/DecompUtil/include/decomp_util/{ellipsoid_decomp.h:96-123, line_segment.h:33-38,57-98,156-252,
decomp_base.h:39-46,83-115}, decomp_geometry/{ellipsoid.h:24-73, polyhedron.h:13-92,114-152, geometric_utils.h:27-35}).

Checked against the reference's own DecompUtil compiled from /root/reference (oracle/decomp_ref.py, tests/test_decomp_cpu.py);
it exists because that library cannot travel to a box without /root/reference being needed to rebuild it.  This restatement follows the
reference statement by statement (same loop order, same strict/non-strict comparisons, epsilon_ = 1e-10 of
decomp_basis/data_type.h:129).  It pins the product's host implementation (faster_b200/csrc/fq_decomp.cpp).
Only tests/ and bench tooling may import it.
"""
import numpy as np

EPS = 1e-10   # decomp_basis/data_type.h:129


def _rx(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[1, 0, 0], [0, c, -s], [0, s, c]])


def _ry(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])


def _rz(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])


def vec3_to_rotation(v):
    """geometric_utils.h:27-35: zero roll, pitch = atan2(-vz, |vxy|), yaw = atan2(vy, vx); R = Rz Ry Rx."""
    pitch = np.arctan2(-v[2], np.hypot(v[0], v[1]))
    yaw = np.arctan2(v[1], v[0])
    return _rz(yaw) @ _ry(pitch) @ _rx(0.0)


def local_bbox_planes(p1, p2, bbox):
    """line_segment.h:57-98 -> list of (point, outward normal) in the reference's order."""
    d = p2 - p1
    dirv = d / np.linalg.norm(d)
    dir_h = np.array([dirv[1], -dirv[0], 0.0])
    if np.linalg.norm(dir_h) == 0:
        dir_h = np.array([-1.0, 0.0, 0.0])
    dir_h = dir_h / np.linalg.norm(dir_h)
    dir_v = np.array([dirv[1] * dir_h[2] - dirv[2] * dir_h[1], dirv[2] * dir_h[0] - dirv[0] * dir_h[2],
                      dirv[0] * dir_h[1] - dirv[1] * dir_h[0]])
    return [(p1 + dir_h * bbox[1], dir_h), (p1 - dir_h * bbox[1], -dir_h),
            (p2 + dirv * bbox[0], dirv), (p1 - dirv * bbox[0], -dirv),
            (p1 + dir_v * bbox[2], dir_v), (p1 - dir_v * bbox[2], -dir_v)]


class _Ellipsoid:
    def __init__(self, C, d):
        self.C, self.d = C, d

    def dist(self, pts):                       # ellipsoid.h:24-27
        Ci = np.linalg.inv(self.C)
        return np.linalg.norm((pts - self.d) @ Ci.T, axis=-1)

    def closest_point(self, pts):              # ellipsoid.h:46-60 (first strict minimum)
        return pts[int(np.argmin(self.dist(pts)))]


def decompose_segment(p1, p2, obs, bbox=(2.0, 2.0, 1.0), inflate=0.42):
    """LineSegment3D: set_obs + dilate(0).  Returns (planes [(point, normal)...], ellipsoid (C, d))."""
    p1 = np.asarray(p1, float)
    p2 = np.asarray(p2, float)
    obs = np.asarray(obs, float).reshape(-1, 3)
    planes_bbox = local_bbox_planes(p1, p2, bbox)
    # set_obs (decomp_base.h:39-46): keep points inside the local bbox, non-exclusive with epsilon_ (polyhedron.h:65-76)
    keep = np.ones(len(obs), bool)
    for pt, n in planes_bbox:
        keep &= ~((obs - pt) @ n > EPS)
    O = obs[keep].copy()
    # find_ellipsoid(0)  (line_segment.h:156-252)
    f = np.linalg.norm(p1 - p2) / 2
    axes = np.array([f, f, f])
    Ri = vec3_to_rotation(p2 - p1)
    d = (p1 + p2) / 2
    E = _Ellipsoid(Ri @ (f * np.eye(3)) @ Ri.T, d)
    Rf = Ri
    if len(O):                                 # obstacle inflation (:178-190), in place
        P = (O - d) @ Ri                       # Ri^T (it - d)
        P = P - np.sign(P) * inflate
        O = P @ Ri.T + d
    inside0 = O[E.dist(O) <= 1] if len(O) else O
    cur = inside0
    while len(cur):
        pw = E.closest_point(cur)
        p = Ri.T @ (pw - d)
        roll = np.arctan2(p[2], p[1])
        Rf = Ri @ _rx(roll)
        p = Rf.T @ (pw - d)
        if p[0] < axes[0]:
            axes[1] = abs(p[1]) / np.sqrt(1 - (p[0] / axes[0]) ** 2)
        E.C = Rf @ np.diag([axes[0], axes[1], axes[1]]) @ Rf.T
        cur = cur[1 - E.dist(cur) > EPS]
    E.C = Rf @ np.diag(axes) @ Rf.T
    cur = inside0[E.dist(inside0) <= 1] if len(inside0) else inside0
    while len(cur):
        pw = E.closest_point(cur)
        p = Rf.T @ (pw - d)
        dd = 1 - (p[0] / axes[0]) ** 2 - (p[1] / axes[1]) ** 2
        if dd > EPS:
            axes[2] = abs(p[2]) / np.sqrt(dd)
        E.C = Rf @ np.diag(axes) @ Rf.T
        cur = cur[1 - E.dist(cur) > EPS]
    # find_polyhedron (decomp_base.h:83-115)
    planes = []
    remain = O
    while len(remain):
        cp = E.closest_point(remain)
        Ci = np.linalg.inv(E.C)
        n = Ci @ Ci.T @ (cp - d)               # ellipsoid.h:65-73
        n = n / np.linalg.norm(n)
        planes.append((cp, n))
        remain = remain[(remain - cp) @ n < 0]
    planes += planes_bbox                      # add_local_bbox (line_segment.h:33-38)
    return planes, (E.C.copy(), d)


def constraints_from_planes(planes, pt_inside, z_ground):
    """LinearConstraint3D(p0, hyperplanes) (polyhedron.h:131-152) + the ground face (jps_manager.cpp:118-122)."""
    A, b = [], []
    for pt, n in planes:
        c = float(pt @ n)
        if n @ pt_inside - c > 0:
            n, c = -n, -c
        A.append(n)
        b.append(c)
    A.append(np.array([0.0, 0.0, -1.0]))
    b.append(-z_ground)
    return np.array(A), np.array(b)


def cvx_ellipsoid_decomp(path, obs, bbox=(2.0, 2.0, 1.0), inflate=0.42, z_ground=0.0):
    """jps_manager.cpp:80-127 -> list of (A, b), one polytope per path segment."""
    path = np.asarray(path, float)
    out = []
    for i in range(len(path) - 1):
        planes, _ = decompose_segment(path[i], path[i + 1], obs, bbox, inflate)
        out.append(constraints_from_planes(planes, (path[i] + path[i + 1]) / 2, z_ground))
    return out
