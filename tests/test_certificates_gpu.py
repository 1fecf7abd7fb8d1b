"""Parity evidence that does NOT go through oracle/fq_oracle.c:
  (1) the CUDA path against HiGHS on the LITERAL model of the reference (oracle/model_fullspace.py: the reference's own
      12N coefficient variables, every row as solverGurobi.cpp writes it), on samples of BASELINE configs 2, 3 and 5;
  (2) for candidates the GPU reports infeasible, the Farkas certificate the solver stopped on (fq_solve_batch_cert) is
      checked on those literal rows: multipliers y >= 0 on the inequality rows, free multipliers on the equality rows
      (initial state, final state, continuity), combination of the rows = 0, combination of the right-hand sides < 0.
      Such a certificate proves infeasibility whatever solver produced it;
  (3) for candidates the GPU reports solved, an OPTIMALITY proof on the same literal rows (oracle/proofs.py): the GPU's
      coefficients satisfy every row, and Karush-Kuhn-Tucker multipliers exist for them (free on the equality rows, >= 0 on
      the tight inequality rows, stationarity residual at rounding level).  For this convex QP that is sufficient: the point
      is the optimum of the reference's model and the reported cost is its minimum -- again whatever solver produced it.
      The prover's own negative controls run on the CPU (tests/test_proofs_cpu.py).
"""
import numpy as np
import pytest

from faster_b200 import corridor as cr

pytestmark = pytest.mark.gpu

CASES = [("cfg2", 10, 3, True, "uav"), ("cfg3", 10, 4, False, "uav"), ("cfg5", 15, 8, True, "ground")]


def _candidates(N, P, ff, profile, seed, n_sig, rng):
    pb = cr.make_corridor(seed, P, N, profile, ff)
    allm = cr.monotone_sigmas(N, P) if P <= 4 else cr.sample_monotone_sigmas(N, P, 400, rng)
    sig = allm[rng.choice(len(allm), n_sig, replace=False)]
    from faster_b200 import capi
    base = max(capi.dt_initial(pb["x0"], pb["xf"], pb["lim"], N), 0.02)
    fac = np.array([1.0, 2.0, 3.0, 5.0, 8.0])
    dts = np.repeat(fac * base, n_sig)
    sigs = np.tile(sig, (len(fac), 1))
    return pb, dts, sigs


@pytest.mark.parametrize("name,N,P,ff,profile", CASES)
def test_cuda_path_against_highs_on_the_literal_model(solver, name, N, P, ff, profile):
    """GPU flag / cost / coefficients vs HiGHS on the literal model.  A feasible flag is additionally PROVED by plugging the
    GPU's coefficients into the literal rows (every equality and inequality of the reference's model holds): that needs no
    solver at all.  HiGHS' QP solver occasionally stops without a verdict on these degenerate problems; such cases count
    only through the row check."""
    from oracle import model_fullspace as mf
    rng = np.random.default_rng(sum(map(ord, name)))
    checked = feas_seen = infeas_seen = highs_agree = 0
    for seed in (6100, 6101, 6102):
        pb, dts, sigs = _candidates(N, P, ff, profile, seed, 6, rng)
        fg, cg, cog, _ = solver.solve_batch(N, pb["x0"], pb["xf"], pb["lim"], pb["polys"], dts, sigs, ff, want_coeffs=True)
        for i in range(len(dts)):
            ok, cost, z, status = mf.solve_highs(N, pb["x0"], pb["xf"], pb["lim"], dts[i], pb["polys"], sigs[i], ff, with_status=True)
            checked += 1
            if fg[i]:
                feas_seen += 1
                assert status != "kInfeasible", (name, seed, i)
                Q, Aeq, beq, Ain, bin_ = mf.build(N, pb["x0"], pb["xf"], pb["lim"], dts[i], pb["polys"], sigs[i], ff)
                zg = cog[i].reshape(-1)
                assert np.abs(Aeq @ zg - beq).max() < 1e-7 and (Ain @ zg - bin_).max() < 1e-7       # feasible: here is the point
                assert abs(float(np.sum((6.0 * cog[i][:, :3]) ** 2)) - cg[i]) <= 1e-9 * max(1.0, cg[i])
                if ok:
                    highs_agree += 1
                    assert abs(cg[i] - cost) <= 1e-5 * max(1.0, abs(cost)), (cg[i], cost)
                    assert cg[i] <= cost * (1 + 1e-6) + 1e-9                                 # never worse than HiGHS' optimum
                    assert np.abs(cog[i] - z).max() <= 1e-3 * max(1.0, np.abs(z).max())     # HiGHS' own accuracy on the coefficients
            else:
                infeas_seen += 1
                assert not ok, (name, seed, i, dts[i], sigs[i])            # proofs of these: the certificate test below
    assert checked == 90 and feas_seen > 10 and infeas_seen > 5 and highs_agree > 5


@pytest.mark.parametrize("name,N,P,ff,profile", CASES)
def test_infeasibility_certificates_hold_on_the_literal_rows(solver, name, N, P, ff, profile):
    from oracle import model_fullspace as mf, proofs
    rng = np.random.default_rng(7 + N + P)
    verified = 0
    for seed in (6200, 6201, 6202, 6203):
        pb, dts, sigs = _candidates(N, P, ff, profile, seed, 8, rng)
        fg, _, cert = solver.solve_batch_cert(N, pb["x0"], pb["xf"], pb["lim"], pb["polys"], dts, sigs, ff)
        f2, _, _, _ = solver.solve_batch(N, pb["x0"], pb["xf"], pb["lim"], pb["polys"], dts, sigs, ff)
        assert np.array_equal(fg, f2)                          # generic (certifying) and specialised kernel agree
        for i in np.flatnonzero(fg == 0)[:12]:
            model = mf.build(N, pb["x0"], pb["xf"], pb["lim"], dts[i], pb["polys"], sigs[i], ff)
            proofs.assert_infeasible(model, N, pb["polys"], sigs[i], cert[i])
            verified += 1
    assert verified >= 10


@pytest.mark.parametrize("name,N,P,ff,profile", CASES)
def test_solved_flags_carry_optimality_proofs(solver, name, N, P, ff, profile):
    """Every candidate the product kernel reports solved: its coefficients are feasible on the literal rows AND carry KKT
    multipliers there (stationarity residual <= 1e-6 relative, complementarity gap <= 1e-7 relative), the reported cost is
    the cost of that point.  Together with the Farkas test above every flag of these batches is proved, not compared."""
    from oracle import model_fullspace as mf, proofs
    rng = np.random.default_rng(11 + N + P)
    proved = active = 0
    worst = 0.0
    for seed in (6400, 6401, 6402):
        pb, dts, sigs = _candidates(N, P, ff, profile, seed, 8, rng)
        fg, cg, cog, _ = solver.solve_batch(N, pb["x0"], pb["xf"], pb["lim"], pb["polys"], dts, sigs, ff, want_coeffs=True)
        for i in np.flatnonzero(fg):
            model = mf.build(N, pb["x0"], pb["xf"], pb["lim"], dts[i], pb["polys"], sigs[i], ff)
            r = proofs.assert_optimal(model, cog[i], cg[i])
            proved += 1
            active += r["n_active"] > 0
            worst = max(worst, r["resid"])
    assert proved >= 20 and active >= 5, (proved, active)
    print("%s: %d optima proved on the literal model, %d with tight rows, worst stationarity residual %.1e" % (name, proved, active, worst))
