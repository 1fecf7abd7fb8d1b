"""The optimality prover (oracle/proofs.py) itself, on the CPU: it accepts the optima of the CPU restatement and of the
tuned CPU port on the literal model of the reference, and it REJECTS points that are feasible but not optimal (negative
controls) -- so that the GPU test built on it (tests/test_certificates_gpu.py) means something.  No solver is trusted:
the multipliers are a certificate, their check is a handful of matrix products."""
import numpy as np
import pytest

from faster_b200 import corridor as cr
from oracle import model_fullspace as mf, proofs

CASES = [("cfg2", 10, 3, True, "uav"), ("cfg3", 10, 4, False, "uav"), ("cfg5", 15, 8, True, "ground")]


def _batch(oracle, N, P, ff, profile, seed, rng, n_sig=5):
    pb = cr.make_corridor(seed, P, N, profile, ff)
    allm = cr.monotone_sigmas(N, P) if P <= 4 else cr.sample_monotone_sigmas(N, P, 400, rng)
    sig = allm[rng.choice(len(allm), n_sig, replace=False)]
    base = max(oracle.dt_initial(pb["x0"], pb["xf"], pb["lim"], N), 0.02)
    fac = np.array([1.5, 2.5, 4.0, 8.0])
    return pb, np.repeat(fac * base, n_sig), np.tile(sig, (len(fac), 1))


@pytest.mark.parametrize("name,N,P,ff,profile", CASES)
def test_oracle_optima_carry_kkt_certificates(oracle, name, N, P, ff, profile):
    rng = np.random.default_rng(len(name) + N + P)
    proved = with_active_corridor_rows = 0
    for seed in (6300, 6301):
        pb, dts, sigs = _batch(oracle, N, P, ff, profile, seed, rng)
        f, c, co = oracle.solve_batch(N, pb["x0"], pb["xf"], pb["lim"], pb["polys"], dts, sigs, ff, want_coeffs=True)
        for i in np.flatnonzero(f):
            model = mf.build(N, pb["x0"], pb["xf"], pb["lim"], dts[i], pb["polys"], sigs[i], ff)
            r = proofs.assert_optimal(model, co[i], c[i])
            proved += 1
            with_active_corridor_rows += r["n_active"] > 0
    assert proved >= 8 and with_active_corridor_rows >= 3


def test_tuned_port_optima_carry_kkt_certificates(oracle):
    """bench.py's CPU arm (oracle/fq_cpu_port.c) answers the same model: its optima are proved the same way."""
    N, P, ff = 10, 3, True
    rng = np.random.default_rng(5)
    pb, dts, sigs = _batch(oracle, N, P, ff, "uav", 6310, rng)
    po_, fo, Ab = oracle.pack_polys(pb["polys"])
    out = oracle.solve_multi_port(N, ff, np.asarray(pb["x0"], float).reshape(1, 9), np.asarray(pb["xf"], float).reshape(1, 9),
                                  np.asarray(pb["lim"], float).reshape(1, 3), np.array([0, P], np.int32), fo, Ab,
                                  np.array([0, len(dts)], np.int32), dts, sigs, threads=2, want_coeffs=True)
    f, c, co = out[0], out[1], out[2]
    assert f.sum() >= 4
    for i in np.flatnonzero(f):
        model = mf.build(N, pb["x0"], pb["xf"], pb["lim"], dts[i], pb["polys"], sigs[i], ff)
        proofs.assert_optimal(model, co[i].reshape(N, 12), c[i])


def test_prover_rejects_feasible_points_that_are_not_optimal(oracle, demo_corridor):
    """Negative controls.  (a) The optimum of a candidate is still FEASIBLE for the same model at a looser corridor (one
    polytope row pushed outwards where it was tight) but no longer optimal there: the fit finds no multipliers.  (b) A
    convex combination of two candidates' optima that share every row (same dt, same assignment, different final state
    is not allowed -- so: the optimum blended with a feasible non-optimal point) is feasible and is rejected too."""
    fx = demo_corridor
    N, dt, sigma = fx["N"], 0.8, [0, 0, 0, 0, 1, 1, 1, 2, 2, 2]
    rc, c, co, _ = oracle.solve_fixed(N, fx["x0"], fx["xf"], fx["lim"], dt, fx["polys"], sigma)
    assert rc == 1
    model = mf.build(N, fx["x0"], fx["xf"], fx["lim"], dt, fx["polys"], sigma)
    r = proofs.assert_optimal(model, co, c)
    assert r["n_active"] >= 1, "the demo candidate should touch its corridor or a box"
    Q, Aeq, beq, Ain, bin_ = model
    z = co.reshape(-1)
    slack = bin_ - Ain @ z
    tight = np.flatnonzero(slack <= 1e-7)
    # (a) drop the tight rows: z stays feasible, the unconstrained-er optimum is elsewhere
    keep = np.setdiff1d(np.arange(len(bin_)), tight)
    ra = proofs.prove_optimal(Q, Aeq, beq, Ain[keep], bin_[keep], z)
    assert ra["eq"] <= 1e-9 and ra["ineq"] <= 1e-9 and ra["resid"] > 1e-4, ra
    with pytest.raises(AssertionError):
        proofs.assert_optimal((Q, Aeq, beq, Ain[keep], bin_[keep]), co)
    # (b) blend with the optimum of the relaxed problem (also feasible for it): feasible, not optimal
    ok, ch, zh = mf.solve_highs(N, fx["x0"], fx["xf"], fx["lim"], dt, [], sigma)      # no corridor rows at all
    assert ok and ch < c
    blend = 0.5 * z + 0.5 * zh.reshape(-1)
    box_rows = 18 * N                                                               # the |v|,|a|,|j| rows come first
    rb = proofs.prove_optimal(Q, Aeq, beq, Ain[:box_rows], bin_[:box_rows], blend)
    assert rb["eq"] <= 1e-7 and rb["ineq"] <= 1e-7 and rb["resid"] > 1e-4, rb
    # and the reported cost is checked against the point
    with pytest.raises(AssertionError):
        proofs.assert_optimal(model, co, c * (1 + 1e-6))
