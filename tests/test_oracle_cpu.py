"""CPU suite: pins the oracle (CPU restatement) and the host-side product logic.  No GPU needed.

PARITY UNPINNED w.r.t. Gurobi (closed source, absent; the reference holds no golden outputs for this path --
SURVEY.md section 8c).  What pins the oracle instead:
  * the literal full-space model of the reference (oracle/model_fullspace.py: 12N coefficient variables, every row
    written as solverGurobi.cpp writes it, no elimination) solved by an independent solver (HiGHS);
  * the closed form of BASELINE config 1 (N=3 whole: zero degrees of freedom);
  * exhaustive enumeration of all P^N assignments for the MIQP minimum on small cases.
"""
import itertools
import warnings

import numpy as np
import pytest

from faster_b200 import capi, corridor as cr
import kernel_mirror as km
from oracle import model_fullspace as mf

warnings.filterwarnings("ignore")


def test_oracle_vs_highs_demo_corridor(oracle, demo_corridor):
    fx = demo_corridor
    N = fx["N"]
    n_feas = 0
    for dt in (0.5, 0.7, 0.8, 1.0):
        for sigma in ([0, 0, 0, 0, 1, 1, 1, 2, 2, 2], [0, 0, 0, 1, 1, 1, 1, 2, 2, 2], [0, 0, 0, 0, 0, 0, 2, 2, 2, 2]):
            rc, c, co, _ = oracle.solve_fixed(N, fx["x0"], fx["xf"], fx["lim"], dt, fx["polys"], sigma)
            ok, ch, coh = mf.solve_highs(N, fx["x0"], fx["xf"], fx["lim"], dt, fx["polys"], sigma)
            assert (rc == 1) == ok
            if ok:
                n_feas += 1
                assert abs(c - ch) <= 1e-6 * max(1.0, ch)
                assert np.abs(co - coh).max() <= 2e-5
    assert n_feas >= 5


@pytest.mark.parametrize("N,P,ff", [(10, 3, True), (10, 4, False), (6, 3, True), (6, 2, False), (15, 8, True)])
def test_oracle_vs_highs_random(oracle, N, P, ff):
    rng = np.random.default_rng(N + P)
    sig_all = cr.monotone_sigmas(N, P) if P <= 4 else cr.sample_monotone_sigmas(N, P, 64, rng)
    checked = feas = 0
    for seed in range(3):
        pb = cr.make_corridor(500 + seed, P, N, "uav" if N < 15 else "ground", ff)
        dti = capi.dt_initial(pb["x0"], pb["xf"], pb["lim"], N)
        for f in (1.5, 2.5, 4.0):
            dt = f * dti
            for s in sig_all[rng.choice(len(sig_all), 3, replace=False)]:
                rc, c, co, _ = oracle.solve_fixed(N, pb["x0"], pb["xf"], pb["lim"], dt, pb["polys"], s, ff)
                ok, ch, coh = mf.solve_highs(N, pb["x0"], pb["xf"], pb["lim"], dt, pb["polys"], s, ff)
                checked += 1
                if rc == 1 and not ok:
                    # HiGHS declares borderline problems infeasible with its own 1e-7 tolerances; require that the
                    # oracle's point really satisfies every literal row
                    _, Aeq, beq, Ain, bin_ = mf.build(N, pb["x0"], pb["xf"], pb["lim"], dt, pb["polys"], s, ff)
                    z = co.reshape(-1)
                    assert np.abs(Aeq @ z - beq).max() < 1e-7 and (Ain @ z - bin_).max() < 1e-7
                    continue
                assert (rc == 1) == ok, (seed, f, s)
                if ok:
                    feas += 1
                    assert abs(c - ch) <= 1e-5 * max(1.0, ch)
                    assert np.abs(co - coh).max() <= 1e-4 * max(1.0, np.abs(coh).max())
    assert checked == 27 and feas >= 3


def test_config1_closed_form(oracle):
    """N=3 whole, no polytopes: the 9 equalities leave no freedom; the jerks are the unique solution of a 3x3 system
    per axis, and the candidate is feasible iff that spline respects the boxes."""
    lim = np.array([5.0, 5.0, 8.0])
    N = 3
    for trial in range(20):
        rng = np.random.default_rng(trial)
        x0 = np.zeros(9)
        xf = np.zeros(9)
        xf[:3] = rng.uniform(1, 4, 3) * rng.choice([-1, 1], 3)
        x0[3:6] = rng.uniform(-1, 1, 3)
        dt = rng.uniform(0.3, 1.5)
        # closed form: final state = Phi^3 s0 + [Phi^2 G, Phi G, G] u
        Phi = np.array([[1, dt, dt * dt / 2], [0, 1, dt], [0, 0, 1]])
        G = np.array([dt ** 3 / 6, dt ** 2 / 2, dt])
        M = np.column_stack([Phi @ Phi @ G, Phi @ G, G])
        feas, cost = True, 0.0
        for ax in range(3):
            s0 = np.array([x0[ax], x0[3 + ax], x0[6 + ax]])
            u = np.linalg.solve(M, np.array([xf[ax], 0.0, 0.0]) - Phi @ Phi @ Phi @ s0)
            cost += float(u @ u)
            s = s0.copy()
            for t in range(3):
                if abs(s[1]) > lim[0] + 1e-9 or abs(s[2]) > lim[1] + 1e-9 or abs(u[t]) > lim[2] + 1e-9:
                    feas = False
                s = Phi @ s + G * u[t]
        rc, c, co, _ = oracle.solve_fixed(N, x0, xf, lim, dt, [], None, True)
        assert (rc == 1) == feas, trial
        if feas:
            assert abs(c - cost) <= 1e-9 * max(1.0, cost)
            assert np.allclose(6 * co[:, 0:3].sum(axis=0) * 0 + co[0, 9:12], x0[:3])


def test_miqp_branch_and_bound_equals_exhaustive(oracle):
    """fqo_solve_miqp (B&B over all P^N assignments, no monotonicity assumption) == brute force on small cases, and
    the monotone list contains the optimum on corridor-shaped inputs."""
    for (N, P, ff, seed) in ((5, 3, True, 11), (6, 2, False, 12), (6, 3, True, 13), (5, 3, False, 14)):
        pb = cr.make_corridor(seed, P, N, "uav", ff)
        dti = capi.dt_initial(pb["x0"], pb["xf"], pb["lim"], N)
        for f in (2.0, 3.5):
            dt = f * dti
            allsig = np.array(list(itertools.product(range(P), repeat=N)), np.uint8)
            feas, cost, _ = oracle.solve_batch(N, pb["x0"], pb["xf"], pb["lim"], pb["polys"], np.full(len(allsig), dt),
                                               allsig, ff, threads=8)
            rc, c, co, sig, nodes = oracle.solve_miqp(N, pb["x0"], pb["xf"], pb["lim"], dt, pb["polys"], ff)
            assert (rc == 1) == bool(feas.any())
            if rc == 1:
                assert abs(c - cost.min()) <= 1e-9 * max(1.0, c)
                mono = cr.monotone_sigmas(N, P)
                fm, cm, _ = oracle.solve_batch(N, pb["x0"], pb["xf"], pb["lim"], pb["polys"], np.full(len(mono), dt),
                                               mono, ff)
                assert fm.any() and abs(cm.min() - c) <= 1e-9 * max(1.0, c)
                assert nodes < len(allsig)


def _dt_initial_numpy(x0, xf, lim, N):
    """getDTInitial restated with numpy's companion-matrix roots (what Eigen's PolynomialSolver does) and float32
    temporaries where the reference declares `float` (solverGurobi.cpp:659-759)."""
    f32 = np.float32
    ts = []
    for i in range(3):
        dp = xf[i] - x0[i]
        ts.append(f32(abs(dp) / lim[0]))
        jerk = f32(np.copysign(1.0, dp) * lim[2])
        accel = f32(np.copysign(1.0, dp) * lim[1])
        a0, v0 = f32(x0[6 + i]), f32(x0[3 + i])

        def minpos(coeffs_low_to_high):
            r = np.roots(np.array(coeffs_low_to_high[::-1], float))
            r = sorted(x.real for x in r if abs(x.imag) < 1e-12)
            for x in r:
                if x > 0:
                    return x
            return 0.0
        ts.append(f32(minpos([x0[i] - xf[i], float(v0), float(a0) / 2.0, float(jerk) / 6.0])))
        ts.append(f32(minpos([x0[i] - xf[i], float(v0), 0.5 * float(accel)])))
    d = float(f32(max(ts)) / f32(N))
    return 0.0 if d > 10000 else d


def test_dt_initial_matches_companion_matrix_restatement(oracle):
    rng = np.random.default_rng(5)
    for k in range(300):
        lim = [5.0, 5.0, 8.0] if k % 2 else [1.4, 1.4, 5.0]
        x0 = np.concatenate([rng.uniform(-5, 5, 3), rng.uniform(-2, 2, 3), rng.uniform(-1, 1, 3)])
        xf = np.concatenate([x0[:3] + rng.uniform(-4, 4, 3), np.zeros(6)])
        if k % 7 == 0:
            xf[0] = x0[0]                        # zero displacement on one axis
        N = int(rng.integers(3, 16))
        a = capi.dt_initial(x0, xf, lim, N)
        b = oracle.dt_initial(x0, xf, lim, N)
        c = _dt_initial_numpy(x0, xf, lim, N)
        assert a == b
        assert abs(a - c) <= 2e-7 * max(1.0, abs(c)), (k, a, c)    # one float32 ulp where a root sits on a rounding edge


def test_fill_x_and_num_samples(oracle, demo_corridor):
    fx = demo_corridor
    rc, c, co, _ = oracle.solve_fixed(fx["N"], fx["x0"], fx["xf"], fx["lim"], 0.8, fx["polys"], [0] * 6 + [2] * 4)
    assert rc == 1
    DC = 0.01
    Xo = oracle.fill_x(fx["N"], co, 0.8, DC)
    Xp = capi.fill_x(fx["N"], co, 0.8, DC)
    assert Xo.shape == Xp.shape == (800, 12) and np.allclose(Xo, Xp, rtol=1e-12, atol=1e-12)   # FMA contraction differs
    assert capi.num_samples(10, 0.8, 0.01) == 800 and capi.num_samples(3, 0.001, 0.01) == 2
    # sample i is at t=(i+1)*DC; the last sample's vel/accel/jerk are forced to zero (solverGurobi.cpp:165-167)
    assert np.all(Xp[-1, 3:] == 0) and np.allclose(Xp[-1, :3], fx["xf"][:3], atol=1e-6)
    t = 0.37
    i = int(round(t / DC)) - 1
    seg = int(t / 0.8)
    tau = (i + 1) * DC - seg * 0.8
    a, b, cc, d = co[seg, 0:3], co[seg, 3:6], co[seg, 6:9], co[seg, 9:12]
    assert np.allclose(Xp[i, :3], a * tau ** 3 + b * tau ** 2 + cc * tau + d, atol=1e-9)
    assert np.allclose(Xp[i, 9:12], 6 * a)


def test_monotone_sigmas():
    from math import comb
    for N, P in ((10, 3), (10, 4), (6, 3), (15, 2)):
        s = capi.monotone_sigmas(N, P)
        assert s.shape == (comb(N + P - 1, P - 1), N)
        assert np.all(np.diff(s.astype(int), axis=1) >= 0) and s.max() == P - 1
        assert len({tuple(r) for r in s}) == len(s)
        assert np.array_equal(s, cr.monotone_sigmas(N, P))


@pytest.mark.parametrize("N,P,ff", [(10, 3, True), (10, 4, False), (3, 0, True), (15, 8, False), (5, 2, True)])
def test_kernel_mathematics_mirror_vs_oracle(oracle, N, P, ff):
    """The CUDA kernel's algorithm, mirrored step for step in numpy on the product's own plan tables, against the
    oracle (which works in un-normalised jerk space with literally written rows)."""
    tab = capi.plan_tables(N, ff)
    rng = np.random.default_rng(N * 7 + P)
    sig_all = cr.monotone_sigmas(N, P) if 0 < P <= 4 else (cr.sample_monotone_sigmas(N, P, 32, rng) if P else np.zeros((1, N), np.uint8))
    n = 0
    for seed in range(2):
        pb = cr.make_corridor(800 + seed, max(P, 1), N, "uav" if N < 15 else "ground", ff)
        polys = pb["polys"] if P else []
        dti = capi.dt_initial(pb["x0"], pb["xf"], pb["lim"], N)
        for f in (1.0, 2.0, 3.0, 5.0):
            for s in sig_all[rng.choice(len(sig_all), min(3, len(sig_all)), replace=False)]:
                dt = f * max(dti, 0.02)
                rc, c, co, _ = oracle.solve_fixed(N, pb["x0"], pb["xf"], pb["lim"], dt, polys, s, ff)
                for normalised, thin in ((False, False), (True, False), (True, True)):
                    # generic kernel / specialised pivot rule / specialised kernel's thin Gram-Schmidt factorisation
                    st, cm, com, _ = km.solve(tab, N, pb["x0"], pb["xf"], pb["lim"], dt, polys, s, ff, normalised, thin)
                    assert (rc == 1) == (st == 1)
                    if rc == 1:
                        assert abs(c - cm) <= 1e-8 * max(1.0, c) and np.abs(co - com).max() <= 1e-7 * max(1.0, np.abs(co).max())
                n += 1
    assert n >= 8


def test_oracle_multi_equals_batches(oracle):
    """The multi-problem entry used by bench.py's CPU legs gives what the per-problem entry gives."""
    N, P = 10, 3
    sig = cr.monotone_sigmas(N, P)
    probs = [cr.make_corridor(20 + k, P, N) for k in range(3)]
    poly_ofs, face_ofs, rows, cand_ofs, dts, sigs = [0], [0], [], [0], [], []
    for k, p in enumerate(probs):
        for A, b in p["polys"]:
            rows.append(np.hstack([A, b[:, None]])); face_ofs.append(face_ofs[-1] + len(b))
        poly_ofs.append(poly_ofs[-1] + P)
        dti = capi.dt_initial(p["x0"], p["xf"], p["lim"], N)
        n = 15 + 4 * k
        dts.append((1.5 + 0.5 * np.arange(n)) * dti); sigs.append(sig[(3 * np.arange(n)) % len(sig)])
        cand_ofs.append(cand_ofs[-1] + n)
    f, c = oracle.solve_multi(N, True, [p["x0"] for p in probs], [p["xf"] for p in probs], [p["lim"] for p in probs],
                              poly_ofs, face_ofs, np.vstack(rows), cand_ofs, np.concatenate(dts), np.vstack(sigs), threads=3)
    for k, p in enumerate(probs):
        fo, co_, _ = oracle.solve_batch(N, p["x0"], p["xf"], p["lim"], p["polys"], dts[k], sigs[k], True)
        a, b = cand_ofs[k], cand_ofs[k + 1]
        assert np.array_equal(f[a:b], fo) and np.array_equal(c[a:b], co_)


def test_oracle_reproduces_committed_goldens(oracle, demo_corridor):
    """tests/golden/corridor_continuous_expected.json (generated once, HiGHS-cross-checked) pins the oracle across rounds."""
    import json
    import os
    exp = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "corridor_continuous_expected.json")))
    fx = demo_corridor
    sig = np.array(exp["sigmas"], np.uint8)
    for dt, fe, ce in zip(exp["dts"], exp["feasible"], exp["cost"]):
        f, c, _ = oracle.solve_batch(fx["N"], fx["x0"], fx["xf"], fx["lim"], fx["polys"], np.full(len(sig), dt), sig, True, threads=4)
        assert f.tolist() == fe
        for k, v in enumerate(ce):
            if v is not None:
                assert abs(c[k] - v) <= 1e-9 * max(1.0, v)
    g = oracle.gen_new_traj(fx["N"], fx["x0"], fx["xf"], fx["lim"], fx["polys"], 0.01, 1.0, 10.0, 1.0, None, True)
    e = exp["gen_new_traj"]
    assert g["solved"] == e["solved"] and g["factor"] == e["factor"] and g["trials"] == e["trials"]
    assert abs(g["dt"] - e["dt"]) <= 1e-15 and abs(g["cost"] - e["cost"]) <= 1e-9 * e["cost"]
