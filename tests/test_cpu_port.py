"""The tuned CPU port (oracle/fq_cpu_port.c, bench.py's CPU arm) against the literal restatement (oracle/fq_oracle.c):
same flags, same costs, same coefficients -- on the configurations bench.py times it on."""
import numpy as np
import pytest

from faster_b200 import corridor as cr


def _batch(cfgs, oracle):
    out = []
    for (N, P, ff, profile, seed) in cfgs:
        pb = cr.make_corridor(seed, P, N, profile, ff)
        sig = cr.monotone_sigmas(N, P) if P <= 4 else cr.sample_monotone_sigmas(N, P, 96, np.random.default_rng(seed))
        sig = sig[::max(1, len(sig) // 48)]
        base = max(oracle.dt_initial(pb["x0"], pb["xf"], pb["lim"], N), 0.02)
        dts = np.repeat(np.arange(1.0, 9.0) * base, len(sig))
        sigs = np.tile(sig, (8, 1))
        out.append((pb, dts, sigs))
    return out


@pytest.mark.parametrize("N,P,ff,profile", [(10, 3, True, "uav"), (10, 4, False, "uav"), (6, 3, True, "uav"), (15, 8, True, "ground"),
                                            (16, 4, False, "uav"), (4, 2, True, "uav")])
def test_port_matches_oracle(oracle, N, P, ff, profile):
    for pb, dts, sigs in _batch([(N, P, ff, profile, 9100 + k) for k in range(3)], oracle):
        po_, fo_, Ab = oracle.pack_polys(pb["polys"])
        args = (N, ff, pb["x0"][None], pb["xf"][None], pb["lim"][None], np.array([0, po_], np.int32), fo_, Ab,
                np.array([0, len(dts)], np.int32), dts, sigs)
        fo, co, cfo = oracle.solve_batch(N, pb["x0"], pb["xf"], pb["lim"], pb["polys"], dts, sigs, ff, True, threads=2)
        fp, cp, cfp = oracle.solve_multi_port(*args, 2, want_coeffs=True)
        assert np.array_equal(fo, fp)
        ok = fo.astype(bool)
        if ok.any():
            assert (np.abs(cp[ok] - co[ok]) / np.abs(co[ok])).max() < 1e-9
            assert np.abs(cfp[ok] - cfo[ok]).max() < 1e-7
        assert np.isinf(cp[~ok]).all()


def test_port_thread_counts_and_tolerance(oracle):
    pb, dts, sigs = _batch([(10, 3, True, "uav", 9200)], oracle)[0]
    po_, fo_, Ab = oracle.pack_polys(pb["polys"])
    args = (10, True, pb["x0"][None], pb["xf"][None], pb["lim"][None], np.array([0, po_], np.int32), fo_, Ab,
            np.array([0, len(dts)], np.int32), dts, sigs)
    base = oracle.solve_multi_port(*args, 1)
    for th in (2, 5, 3):                                   # pool resized between calls
        f, c = oracle.solve_multi_port(*args, th)
        assert np.array_equal(f, base[0]) and np.array_equal(c, base[1])
    # the run-time row tolerance is mirrored
    oracle.port_lib().fqc_set_row_tol(1e-3)
    oracle.set_row_tol(1e-3)
    try:
        f3, _ = oracle.solve_multi_port(*args, 2)
        fo3, _, _ = oracle.solve_batch(10, pb["x0"], pb["xf"], pb["lim"], pb["polys"], dts, sigs, True, False, threads=2)
        assert np.array_equal(f3, fo3)
        assert f3.sum() >= base[0].sum()
    finally:
        oracle.port_lib().fqc_set_row_tol(1e-8)
        oracle.set_row_tol(1e-8)
