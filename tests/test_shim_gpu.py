"""The drop-in C++ class include/solverGurobi.hpp, driven the way Faster::replan() drives the reference class
(tests/cpp/shim_driver.cpp), against the oracle's sequential sweep with branch-and-bound over ALL assignments."""
import json
import os
import subprocess

import numpy as np
import pytest

from faster_b200 import corridor as cr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def driver(built_lib, tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("shim") / "shim_driver")
    libdir = os.path.dirname(built_lib)
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "shim_driver.cpp"), "-o", exe, "-L", libdir,
                           "-lfaster_b200", "-Wl,-rpath," + libdir])
    return exe


def _run(exe, N, ff, DC, lim, fi, fl, finc, x0, xf, polys, mode=None):
    toks = [N, int(ff), DC, *lim, fi, fl, finc, *x0, *xf, len(polys)]
    for A, b in polys:
        toks.append(len(b))
        for f in range(len(b)):
            toks += [*A[f], b[f]]
    if mode is not None:
        toks.append(int(mode))
    out = subprocess.run([exe], input=" ".join(repr(float(t)) if isinstance(t, (float, np.floating)) else str(t) for t in toks),
                         capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    return json.loads(out.stdout.strip().splitlines()[-1])


def test_shim_builds_without_gpu(driver):
    assert os.path.exists(driver)


@pytest.mark.gpu
def test_shim_matches_oracle_sweep(driver, oracle, demo_corridor):
    fx = demo_corridor
    cases = [(fx["N"], True, fx["x0"], fx["xf"], fx["lim"], fx["polys"])]
    for seed in range(3):
        pb = cr.make_corridor(600 + seed, 3, 6)                       # shipped yaml sizes: N=6, <=3 polytopes
        cases.append((6, True, pb["x0"], pb["xf"], pb["lim"], pb["polys"]))
        pb = cr.make_corridor(700 + seed, 3, 6, force_final=False)
        cases.append((6, False, pb["x0"], pb["xf"], pb["lim"], pb["polys"]))
    pb = cr.make_corridor(5, 3, 10)                                   # infeasible corridor: sweep exhausts
    cases.append((10, True, pb["x0"], pb["xf"], pb["lim"], pb["polys"]))
    DC = 0.01
    n_solved = 0
    for N, ff, x0, xf, lim, polys in cases:
        r = _run(driver, N, ff, DC, lim, 1.0, 10.0, 1.0, x0, xf, polys)
        o = oracle.gen_new_traj(N, x0, xf, lim, polys, DC, 1.0, 10.0, 1.0, None, ff)
        assert bool(r["solved"]) == o["solved"]
        assert r["trials"] == o["trials"]
        assert r["dt"] == o["dt"]                                     # identical dt choice, bit for bit
        if o["solved"]:
            n_solved += 1
            assert r["factor"] == o["factor"]
            assert abs(r["cost"] - o["cost"]) <= 1e-7 * max(1.0, o["cost"])
            co = np.array(r["coeffs"]).reshape(N, 12)
            assert np.abs(co - o["coeffs"]).max() <= 1e-6 * max(1.0, np.abs(o["coeffs"]).max())
            X = oracle.fill_x(N, o["coeffs"], o["dt"], DC)
            assert r["n_samples"] == len(X)
            assert np.allclose(r["x_first"], X[0, :6], atol=1e-6)
            assert np.allclose(r["x_mid"], X[len(X) // 2, :3], atol=1e-6)
            assert np.allclose(r["x_last"][:3], X[-1, :3], atol=1e-6) and r["x_last"][3] == 0.0
            assert r["second_solved"] == 1 and r["second_factor"] == o["factor"]
    assert n_solved >= 4


@pytest.fixture(scope="module")
def replan_driver(built_lib, tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("replan") / "replan_driver")
    libdir = os.path.dirname(built_lib)
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "replan_driver.cpp"), "-o", exe, "-L", libdir,
                           "-lfaster_b200", "-Wl,-rpath," + libdir])
    return exe


def _replay_with_oracle(oracle, N, DC, lim, gamma, verts, obs, n_rep):
    """The rules of tests/cpp/replan_driver.cpp replayed with the oracle (B&B over all assignments, numpy decomposition)."""
    from oracle import decomp_oracle as do
    out = []
    A = np.concatenate([verts[0], np.zeros(6)])
    seg0 = 0
    win_w, win_s = (1.0, 10.0), (1.0, 10.0)
    for rep in range(n_rep):
        wpath = [A[:3]] + [verts[k] for k in range(seg0 + 1, len(verts))][:3]
        if len(wpath) < 2:
            break
        wpath = np.array(wpath)
        xf = np.concatenate([wpath[-1], np.zeros(6)])
        polys = do.cvx_ellipsoid_decomp(wpath, obs, (2.0, 2.0, 1.0), 0.42, 0.0)
        w = oracle.gen_new_traj(N, A, xf, lim, polys, DC, win_w[0], win_w[1], 1.0, None, True)
        rec = {"whole": w}
        out.append(rec)
        if not w["solved"]:
            break
        Xw = oracle.fill_x(N, w["coeffs"], w["dt"], DC)
        kR = int(0.6 * len(Xw))
        R = Xw[kR]
        n_seg = len(wpath) - 1
        segR = min(int(((kR + 1) * DC) / (w["dt"] * N / float(n_seg))), n_seg - 1)
        spath = np.vstack([R[:3], wpath[segR + 1:]])
        xfs = np.concatenate([spath[-1], np.zeros(6)])
        spolys = do.cvx_ellipsoid_decomp(spath, obs, (2.0, 2.0, 1.0), 0.42, 0.0)
        s = oracle.gen_new_traj(N, R[:9], xfs, lim, spolys, DC, win_s[0], win_s[1], 1.0, None, False)
        rec["safe"] = s
        rec["R"] = R[:3]
        if not s["solved"]:
            break
        win_w = (max(w["factor"] - gamma, 1.0), w["factor"] + gamma)
        win_s = (max(s["factor"] - gamma, 1.0), s["factor"] + gamma)
        kA = int(0.3 * len(Xw))
        A = Xw[kA][:9].copy()
        seg0 += min(int(((kA + 1) * DC) / (w["dt"] * N / float(n_seg))), n_seg - 1)
    return out


@pytest.mark.gpu
def test_replan_sequence_with_drop_in_class(replan_driver, oracle):
    """Several consecutive replans (whole -> safe -> factor-window update, faster.cpp:406-430,:521-537,:582-588) through
    the drop-in C++ class, replayed step by step with the oracle."""
    N, DC, gamma, n_rep = 6, 0.01, 3.0, 4                      # shipped yaml: N_whole = N_safe = 6
    lim = [5.0, 5.0, 8.0]
    compared = 0
    for seed in (300, 301, 303, 305):
        obs, centres, radii = cr.make_forest(seed)
        try:
            verts = cr.forest_path(seed + 1, centres, radii, 6, clearance=0.42 * 1.45 + 0.05)
        except RuntimeError:
            continue
        toks = [N, DC, *lim, gamma, len(verts), *verts.reshape(-1), len(obs), *obs.reshape(-1), n_rep]
        res = subprocess.run([replan_driver], input=" ".join(repr(float(t)) if isinstance(t, (float, np.floating)) else str(t) for t in toks),
                             capture_output=True, text=True, timeout=120)
        assert res.returncode == 0, res.stderr
        got = json.loads(res.stdout.strip().splitlines()[-1])
        exp = _replay_with_oracle(oracle, N, DC, lim, gamma, verts, obs, n_rep)
        assert len(got) == len(exp)
        for g, e in zip(got, exp):
            for kind in ("whole", "safe"):
                if kind not in e:
                    assert kind not in g or not g[kind]["solved"]
                    continue
                assert bool(g[kind]["solved"]) == e[kind]["solved"], (seed, g["rep"], kind)
                assert g[kind]["trials"] == e[kind]["trials"]
                if e[kind]["solved"]:
                    assert g[kind]["factor"] == e[kind]["factor"]
                    assert abs(g[kind]["dt"] - e[kind]["dt"]) <= 1e-12 * e[kind]["dt"]
                    assert abs(g[kind]["cost"] - e[kind]["cost"]) <= 1e-6 * max(1.0, e[kind]["cost"])
                    compared += 1
            if "R" in e:
                assert np.allclose(g["safe"]["x0"], e["R"], atol=1e-7)
    assert compared >= 8


@pytest.mark.gpu
def test_shim_exact_mode_matches_oracle(driver, oracle, demo_corridor):
    """setAssignmentMode(EXACT): branch-and-bound on the GPU behind the drop-in class (N=10, P=3: 59 049 assignments)."""
    fx = demo_corridor
    cases = [(fx["N"], True, fx["x0"], fx["xf"], fx["lim"], fx["polys"])]
    for seed in range(3):
        pb = cr.make_corridor(640 + seed, 4, 10, force_final=False)
        cases.append((10, False, pb["x0"], pb["xf"], pb["lim"], pb["polys"]))
    for N, ff, x0, xf, lim, polys in cases:
        r = _run(driver, N, ff, 0.01, lim, 1.0, 10.0, 1.0, x0, xf, polys, mode=3)
        o = oracle.gen_new_traj(N, x0, xf, lim, polys, 0.01, 1.0, 10.0, 1.0, None, ff)
        assert r["mode"] == 3 and r["exact"] == 1
        assert bool(r["solved"]) == o["solved"] and r["trials"] == o["trials"]
        if o["solved"]:
            assert r["bnb_nodes"] > 0 and r["factor"] == o["factor"]
            assert abs(r["cost"] - o["cost"]) <= 1e-7 * max(1.0, o["cost"])
            co = np.array(r["coeffs"]).reshape(N, 12)
            assert np.abs(co - o["coeffs"]).max() <= 1e-6 * max(1.0, np.abs(o["coeffs"]).max())
            assert len(r["assignment"]) == N
