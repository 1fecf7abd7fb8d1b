"""Host-side convex decomposition (fq_ellipsoid_decomp, faster_b200/csrc/fq_decomp.cpp) against the numpy restatement
of DecompUtil's EllipsoidDecomp3D + JPS_Manager::cvxEllipsoidDecomp (oracle/decomp_oracle.py)."""
import numpy as np
import pytest

from faster_b200 import capi, corridor as cr
from oracle import decomp_oracle as do


def _same(polys_a, polys_b, tol=1e-9):
    """Same faces.  The order of the obstacle-derived faces may differ where two obstacle points are equidistant from
    the ellipsoid to the last bit (regularly sampled cylinders produce such ties; the product evaluates C^-1 in closed
    form, the oracle inverts C numerically), so rows are compared as sorted sets; the six bounding-box faces and the
    ground face must sit at the end in the reference's order."""
    assert len(polys_a) == len(polys_b)
    for (A1, b1), (A2, b2) in zip(polys_a, polys_b):
        assert A1.shape == A2.shape, "face counts differ"
        assert np.abs(A1[-7:] - A2[-7:]).max() <= tol and np.abs(b1[-7:] - b2[-7:]).max() <= tol
        r1 = np.hstack([A1, b1[:, None]])
        r2 = np.hstack([A2, b2[:, None]])
        r1 = r1[np.lexsort(np.round(r1, 7).T[::-1])]
        r2 = r2[np.lexsort(np.round(r2, 7).T[::-1])]
        assert np.abs(r1 - r2).max() <= tol


@pytest.mark.parametrize("seed", range(6))
def test_forest_decomposition_matches_oracle(built_lib, seed):
    obs, centres, radii = cr.make_forest(40 + seed)
    path = cr.forest_path(140 + seed, centres, radii, 4, clearance=0.42 * 1.45 + 0.05)
    a = capi.ellipsoid_decomp(path, obs, (2.0, 2.0, 1.0), 0.42, 0.0)
    b = do.cvx_ellipsoid_decomp(path, obs, (2.0, 2.0, 1.0), 0.42, 0.0)
    _same(a, b)
    for k, (A, bb) in enumerate(a):
        # conventions the solver relies on: unit normals, segment inside, bbox + ground faces last
        assert np.allclose(np.linalg.norm(A, axis=1), 1.0, atol=1e-12)
        for p in (path[k], path[k + 1], 0.5 * (path[k] + path[k + 1])):
            assert (A @ p - bb).max() <= 1e-9
        assert np.array_equal(A[-1], [0.0, 0.0, -1.0]) and bb[-1] == 0.0
        assert len(bb) >= 7
        # no (inflated) obstacle point is strictly inside the polytope: every kept point is cut by some face
        inside = np.all(obs @ A.T - bb < -0.42 * 1.8, axis=1)
        assert not inside.any()


def test_edge_inputs(built_lib):
    path = np.array([[0.0, 0.0, 1.0], [1.2, 0.3, 1.1]])
    # no obstacles at all: only the six bbox faces and the ground face
    a = capi.ellipsoid_decomp(path, np.zeros((0, 3)))
    b = do.cvx_ellipsoid_decomp(path, np.zeros((0, 3)))
    _same(a, b)
    assert len(a[0][1]) == 7
    # a vertical segment (dir_h degenerates, line_segment.h:64-70) and obstacles exactly on a bbox face
    path = np.array([[0.0, 0.0, 0.5], [0.0, 0.0, 1.6]])
    obs = np.array([[1.0, 0.2, 1.0], [-0.8, -0.9, 0.7], [2.0, 0.0, 1.0], [0.3, 1.1, 1.4], [5.0, 5.0, 5.0]])
    _same(capi.ellipsoid_decomp(path, obs, inflate=0.2), do.cvx_ellipsoid_decomp(path, obs, inflate=0.2))
    # capacity too small -> error code, not a crash
    with pytest.raises(capi.FqError):
        capi.ellipsoid_decomp(np.array([[0.0, 0, 1], [1, 0, 1]]), np.zeros((0, 3)), cap_rows=3)


def test_random_clouds(built_lib):
    rng = np.random.default_rng(9)
    for k in range(25):
        p1 = rng.uniform(-2, 2, 3)
        p2 = p1 + rng.uniform(-1.5, 1.5, 3)
        if np.linalg.norm(p2 - p1) < 0.3:
            continue
        mid = 0.5 * (p1 + p2)
        obs = mid + rng.normal(size=(int(rng.integers(1, 200)), 3)) * rng.uniform(0.5, 2.5)
        d = np.linalg.norm(np.cross(obs - p1, obs - p2), axis=1) / np.linalg.norm(p2 - p1)
        obs = obs[d > 0.6]                       # keep the segment itself obstacle-free
        _same(capi.ellipsoid_decomp(np.array([p1, p2]), obs, inflate=0.2), do.cvx_ellipsoid_decomp(np.array([p1, p2]), obs, inflate=0.2))
