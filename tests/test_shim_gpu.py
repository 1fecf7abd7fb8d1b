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


def _run(exe, N, ff, DC, lim, fi, fl, finc, x0, xf, polys):
    toks = [N, int(ff), DC, *lim, fi, fl, finc, *x0, *xf, len(polys)]
    for A, b in polys:
        toks.append(len(b))
        for f in range(len(b)):
            toks += [*A[f], b[f]]
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
