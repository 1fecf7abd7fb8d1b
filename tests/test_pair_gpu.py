"""Chained replan (fq_replan_pairs*: whole sweep -> R -> safe sweep in one submission) and the multi-GPU entry points,
against the CPU restatement of the same chain (oracle/pair_oracle.py, following faster.cpp:406-430,:474-475,:521-537)."""
import ctypes as C
import os

import numpy as np
import pytest

from faster_b200 import capi, corridor as cr

pytestmark = pytest.mark.gpu


def _synthetic_pairs(n, seed0, n_fac=6, n_sig=16, N=10):
    """whole: synthetic corridor with 3 polytopes; safe: a 4-polytope corridor around the same start (its x0 is replaced
    by R on the device, so only its polytopes / xf matter)."""
    whole = [cr.make_corridor(seed0 + j, 3, N, "uav", True) for j in range(n)]
    safe = []
    for j, w in enumerate(whole):
        s = cr.make_corridor(seed0 + j, 4, N, "uav", False)          # same seed: the first 3 vertices coincide
        safe.append(s)
    sw = cr.monotone_sigmas(N, 3)
    ss = cr.monotone_sigmas(N, 4)
    sw = sw[np.linspace(0, len(sw) - 1, n_sig).round().astype(int)]
    ss = ss[np.linspace(0, len(ss) - 1, n_sig).round().astype(int)]
    fac = np.arange(1.0, n_fac + 1)
    return capi.make_pair_workload(whole, safe, fac, sw, fac, ss, DC=0.01, r_fraction=0.3)


def _check_against_oracle(g, o, n):
    r = g["results"]
    assert np.array_equal(g["feasible_whole"], o["feasible_whole"]), "whole flags differ"
    assert np.array_equal(r["whole_dt_index"], o["whole_dt_index"])
    assert np.array_equal(r["whole_sigma_index"], o["whole_sigma_index"])
    ok = o["whole_dt_index"] >= 0
    assert ok.any()
    assert np.allclose(r["whole_cost"][ok], o["whole_cost"][ok], rtol=1e-8)
    assert np.array_equal(r["whole_dt"][ok], o["whole_dt"][ok]), "whole dt not bit-equal"
    assert np.array_equal(r["n_samples_whole"], o["n_samples_whole"]) and np.array_equal(r["k_safe"], o["k_safe"])
    assert np.abs(r["R"][ok] - o["R"][ok]).max() < 1e-8
    assert np.isnan(r["R"][~ok]).all()
    assert np.abs(g["coeffs_whole"][ok] - o["coeffs_whole"][ok]).max() < 1e-6
    fok = o["feasible_whole"].astype(bool)
    assert np.allclose(g["cost_whole"][fok], o["cost_whole"][fok], rtol=1e-8)
    # safe side (its inputs carry the 1e-12 differences of R: compare values, not bits)
    assert np.array_equal(g["feasible_safe"], o["feasible_safe"]), "safe flags differ"
    assert np.array_equal(r["safe_dt_index"], o["safe_dt_index"])
    sok = o["safe_dt_index"] >= 0
    assert np.allclose(r["safe_cost"][sok], o["safe_cost"][sok], rtol=1e-6)
    assert np.abs(g["coeffs_safe"][sok] - o["coeffs_safe"][sok]).max() < 1e-5
    assert np.allclose(r["safe_dt_base"][ok], o["safe_dt_base_own"][ok], rtol=1e-6)


def test_pairs_match_oracle_chain(solver, oracle):
    from oracle import pair_oracle
    w = _synthetic_pairs(12, 7100)
    g = solver.replan_pairs(w)
    r = g["results"]
    # the oracle chain runs on the device's own time-allocation bases: a last-bit difference in getDTInitial (checked
    # separately below) must not disturb the comparison of the solves
    o = pair_oracle.replan_pairs(w, threads=4, dt_base_whole=r["whole_dt_base"], dt_base_safe=r["safe_dt_base"])
    _check_against_oracle(g, o, 12)
    assert np.array_equal(r["whole_dt_base"], o["whole_dt_base_own"]), "device getDTInitial differs from the host's"
    # a second call on the same context (scratch reuse) gives the same bits
    g2 = solver.replan_pairs(w)
    assert g2["results"].tobytes() == g["results"].tobytes()
    assert np.array_equal(g2["cost_safe"], g["cost_safe"])


def test_device_dt_initial_and_fillx_match_host(solver):
    """getDTInitial (solverGurobi.cpp:659-759) and the sample R of fillX computed on the device vs the host code, on
    problems WITHOUT polytopes (no corridor rows: the sweep is decided by the boxes alone) -- thousands of states."""
    rng = np.random.default_rng(5)
    n, N = 3000, 8
    x0 = np.zeros((n, 9)); xfw = np.zeros((n, 9)); xfs = np.zeros((n, 9))
    x0[:, :3] = rng.uniform(-5, 5, (n, 3)); x0[:, 3:6] = rng.uniform(-3, 3, (n, 3)); x0[:, 6:9] = rng.uniform(-2, 2, (n, 3))
    xfw[:, :3] = x0[:, :3] + rng.uniform(-4, 4, (n, 3))
    xfs[:, :3] = x0[:, :3] + rng.uniform(-4, 4, (n, 3))
    x0[:50, 3:] = 0.0                                     # rest-to-rest cases (roots of the reduced polynomials)
    lim = np.tile(np.array([5.0, 5.0, 8.0]), (n, 1))
    zero = np.zeros(n + 1, np.int32)
    fac = np.arange(1.0, 13.0)
    w = dict(n_prob=n, N_whole=N, N_safe=N, DC=0.01, r_fraction=0.4, x0=x0, xf_whole=xfw, xf_safe=xfs, lim=lim,
             poly_ofs_whole=zero, face_ofs_whole=np.zeros(1, np.int32), Ab_whole=np.zeros((1, 4)), poly_ofs_safe=zero,
             face_ofs_safe=np.zeros(1, np.int32), Ab_safe=np.zeros((1, 4)), factors_whole=fac, sigmas_whole=np.zeros((1, N), np.uint8),
             factors_safe=fac, sigmas_safe=np.zeros((1, N), np.uint8), max_faces_whole=1, max_poly_faces_whole=0,
             max_faces_safe=1, max_poly_faces_safe=0)
    g = solver.replan_pairs(w)
    r = g["results"]
    host = np.array([max(capi.dt_initial(x0[j], xfw[j], lim[j], N), 0.02) for j in range(n)])
    assert np.array_equal(r["whole_dt_base"], host)
    ok = r["whole_dt_index"] >= 0
    assert ok.sum() > n // 2
    bad = 0
    for j in np.flatnonzero(ok)[:600]:
        X = capi.fill_x(N, g["coeffs_whole"][j], r["whole_dt"][j], 0.01)
        k = min(len(X) - 1, int(0.4 * len(X)))
        assert r["n_samples_whole"][j] == len(X) and r["k_safe"][j] == k
        assert np.array_equal(r["R"][j], X[k, :9])         # same source compiled for host and device: bit-equal
        hs = max(capi.dt_initial(r["R"][j], xfs[j], lim[j], N), 0.02)
        bad += hs != r["safe_dt_base"][j]
    assert bad == 0


def test_pairs_device_pointer_entry(solver):
    torch = pytest.importorskip("torch")
    w = _synthetic_pairs(8, 7300)
    g = solver.replan_pairs(w)
    dev = torch.device("cuda", 0)
    d = {k: torch.from_numpy(np.ascontiguousarray(w[k])).to(dev) for k in capi.PAIR_INPUT_KEYS}
    n = w["n_prob"]
    ncw = n * len(w["factors_whole"]) * len(w["sigmas_whole"]); ncs = n * len(w["factors_safe"]) * len(w["sigmas_safe"])
    out = dict(feasible_whole=torch.zeros(ncw, dtype=torch.uint8, device=dev), cost_whole=torch.zeros(ncw, dtype=torch.float64, device=dev),
               feasible_safe=torch.zeros(ncs, dtype=torch.uint8, device=dev), cost_safe=torch.zeros(ncs, dtype=torch.float64, device=dev),
               results=torch.zeros(n * 144, dtype=torch.uint8, device=dev))
    gathered = torch.zeros(n * 144, dtype=torch.uint8, device=dev)
    a = capi.pair_args(w, lambda k: d[k].data_ptr(), {k: v.data_ptr() for k, v in out.items()})
    st = torch.cuda.Stream(device=dev)
    torch.cuda.synchronize()        # the fills above run on torch's stream
    solver.replan_pairs_dev(a, gathered.data_ptr(), st.cuda_stream)
    solver.replan_pairs_dev(a, gathered.data_ptr(), st.cuda_stream)       # back to back on one stream: ordered
    st.synchronize()
    res = np.frombuffer(out["results"].cpu().numpy().tobytes(), capi.PAIR_RESULT_DTYPE)
    assert res.tobytes() == g["results"].tobytes()
    assert gathered.cpu().numpy().tobytes() == g["results"].tobytes()    # world of one: the gather is a copy
    assert np.array_equal(out["feasible_safe"].cpu().numpy(), g["feasible_safe"])
    assert np.array_equal(out["cost_whole"].cpu().numpy(), g["cost_whole"])


def test_pairs_argument_errors(solver):
    w = _synthetic_pairs(2, 7400)
    bad = dict(w); bad["r_fraction"] = 1.5
    with pytest.raises(capi.FqError):
        solver.replan_pairs(bad)
    bad = dict(w); bad["sigmas_safe"] = w["sigmas_safe"].copy(); bad["sigmas_safe"][0, 0] = 9
    with pytest.raises(capi.FqError):
        solver.replan_pairs(bad)
    bad = dict(w); bad["x0"] = w["x0"].copy(); bad["x0"][1, 2] = np.nan
    with pytest.raises(capi.FqError):
        solver.replan_pairs(bad)


def test_forest_fixture_pairs_match_oracle(solver, oracle):
    """BASELINE config 4's committed corridors (bench_data/cfg4_forest.npz): a slice through the chain vs the oracle."""
    from oracle import pair_oracle
    import bench
    w = bench.load_cfg4(0, 6)
    g = solver.replan_pairs(w)
    r = g["results"]
    o = pair_oracle.replan_pairs(w, threads=8, dt_base_whole=r["whole_dt_base"], dt_base_safe=r["safe_dt_base"])
    _check_against_oracle(g, o, 6)
    assert (r["whole_dt_index"] >= 0).all()              # the fixture keeps only corridors whose whole sweep succeeds
    assert np.abs(r["R"] - w["R_oracle"]).max() < 1e-7


# ---------------------------------------------------------------------------------------------------------------------
# several GPUs behind the ABI
# ---------------------------------------------------------------------------------------------------------------------
def _n_gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("n_gpus", [1, 2])
def test_group_context_shards_pairs(solver, n_gpus):
    if _n_gpus() < n_gpus:
        pytest.skip("needs %d GPUs" % n_gpus)
    w = _synthetic_pairs(7, 7500)                         # 7 corridors: uneven shards
    ref = solver.replan_pairs(w)
    g = capi.Solver(n_gpus=n_gpus, devices=list(range(n_gpus)))
    assert g.comm_info()[1] == n_gpus
    out = g.replan_pairs(w)
    assert out["results"].tobytes() == ref["results"].tobytes()
    for k in ("feasible_whole", "cost_whole", "feasible_safe", "cost_safe", "coeffs_whole", "coeffs_safe"):
        assert np.array_equal(out[k], ref[k]), k
    out2 = g.replan_pairs(w, deferred=True)
    g.wait()
    assert out2["results"].tobytes() == ref["results"].tobytes()
    g.close()


@pytest.mark.parametrize("n_gpus", [1, 2])
def test_solve_multi_sharded_winners(solver, n_gpus):
    if _n_gpus() < n_gpus:
        pytest.skip("needs %d GPUs" % n_gpus)
    N, P = 10, 3
    probs = [cr.make_corridor(7600 + j, P, N, "uav", True) for j in range(5)]
    sig = cr.monotone_sigmas(N, P)[::3]
    x0 = np.array([p["x0"] for p in probs]); xf = np.array([p["xf"] for p in probs]); lim = np.array([p["lim"] for p in probs])
    po_, fo_, rows, co, dts, sgs = [0], [0], [], [0], [], []
    for j, p in enumerate(probs):
        for A, b in p["polys"]:
            rows.append(np.hstack([A, b[:, None]])); fo_.append(fo_[-1] + len(b))
        po_.append(po_[-1] + P)
        nf = 3 + j                                        # different candidate counts per problem
        base = max(capi.dt_initial(p["x0"], p["xf"], p["lim"], N), 0.02)
        dts.append(np.repeat(np.arange(1.0, nf + 1) * base, len(sig))); sgs.append(np.tile(sig, (nf, 1)))
        co.append(co[-1] + nf * len(sig))
    args = (N, True, x0, xf, lim, np.array(po_, np.int32), np.array(fo_, np.int32), np.ascontiguousarray(np.vstack(rows)),
            np.array(co, np.int32), np.concatenate(dts), np.ascontiguousarray(np.vstack(sgs)))
    feas, cost, _, _ = solver.solve_multi(*args)
    g = capi.Solver(n_gpus=n_gpus, devices=list(range(n_gpus)))
    f2, c2, wi, wc = g.solve_multi_sharded(*args)
    g.close()
    assert np.array_equal(f2, feas) and np.array_equal(c2, cost)
    for j in range(5):
        f = feas[co[j]:co[j + 1]].astype(bool); c = cost[co[j]:co[j + 1]]; d = args[9][co[j]:co[j + 1]]
        if not f.any():
            assert wi[j] == -1 and np.isinf(wc[j])
            continue
        dmin = d[f].min()
        cand = np.flatnonzero(f & (d == dmin))
        best = cand[np.argmin(c[cand])]
        assert wi[j] == best and wc[j] == c[best]


def test_two_ranks_one_process_each(tmp_path):
    """One process per GPU (the torchrun shape): fq_comm_init on two ranks, fq_replan_pairs_dev with the all-gather of
    the result records, compared with a single-GPU run of both shards."""
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = os.path.join(root, "tests", "two_rank_worker.py")
    uid = tmp_path / "uid.bin"
    procs = [subprocess.Popen([sys.executable, script, str(r), "2", str(uid), str(tmp_path / ("out%d.npy" % r))], cwd=root)
             for r in range(2)]
    try:
        for p in procs:
            assert p.wait(timeout=300) == 0
    finally:
        for p in procs:                                   # a rank that died leaves its peer waiting inside NCCL: do not leak it
            if p.poll() is None:
                p.kill()
    a = np.load(tmp_path / "out0.npy"); b = np.load(tmp_path / "out1.npy")
    assert a.tobytes() == b.tobytes()                     # every rank holds every corridor's record
    s = capi.Solver(0)
    w0, w1 = _synthetic_pairs(5, 7700), _synthetic_pairs(5, 7700 + 1000)
    r0 = s.replan_pairs(w0, want_candidates=False, want_coeffs=False)["results"]
    r1 = s.replan_pairs(w1, want_candidates=False, want_coeffs=False)["results"]
    s.close()
    assert a.tobytes() == r0.tobytes() + r1.tobytes()


def test_early_exit_keeps_the_winners(solver):
    """Option "sweep_early_exit": candidates that cannot win genNewTraj's selection (a smaller dt already has a feasible
    candidate, solverGurobi.cpp:445-446) are not evaluated; the result records are bit-identical, the per-candidate flags
    only lose feasible entries at larger dt."""
    import bench
    for w in (_synthetic_pairs(9, 7900, n_fac=8, n_sig=32), bench.load_cfg4(64, 16)):
        full = solver.replan_pairs(w)
        s2 = capi.Solver(0)
        s2.set_option("sweep_early_exit", 1)
        ee = s2.replan_pairs(w)
        s2.close()
        assert ee["results"].tobytes() == full["results"].tobytes()
        assert np.array_equal(ee["coeffs_whole"], full["coeffs_whole"]) and np.array_equal(ee["coeffs_safe"], full["coeffs_safe"])
        for k in ("whole", "safe"):
            fe, ff = ee["feasible_" + k].astype(bool), full["feasible_" + k].astype(bool)
            assert not (fe & ~ff).any()                               # nothing becomes feasible
            assert fe.sum() < ff.sum()                                # and work was skipped
            n, nf, ns = w["n_prob"], len(w["factors_" + k]), len(w["sigmas_" + k])
            fe, ff = fe.reshape(n, nf, ns), ff.reshape(n, nf, ns)
            for j in range(n):
                d = ee["results"][k + "_dt_index"][j]
                if d >= 0:                                            # everything up to the winning factor is complete
                    assert np.array_equal(fe[j, :d + 1], ff[j, :d + 1])
                    assert np.array_equal(ee["cost_" + k].reshape(n, nf, ns)[j, :d + 1], full["cost_" + k].reshape(n, nf, ns)[j, :d + 1])


@pytest.mark.parametrize("n_gpus", [1, 2])
def test_cpp_driver_several_gpus_through_the_abi(built_lib, tmp_path, n_gpus):
    """tests/cpp/multi_driver.cpp: one C++ process, fq_create_multi via SolverGurobi::setDevices, fq_replan_pairs on the
    group == a single-GPU context, bit for bit."""
    if _n_gpus() < n_gpus:
        pytest.skip("needs %d GPUs" % n_gpus)
    import json
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "multi_driver")
    libdir = os.path.dirname(built_lib)
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-I", os.path.join(root, "include"), os.path.join(root, "tests", "cpp", "multi_driver.cpp"),
                           "-o", exe, "-L", libdir, "-lfaster_b200", "-Wl,-rpath," + libdir])
    out = subprocess.run([exe, str(n_gpus)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["mismatch"] == 0 and d["world"] == n_gpus and d["whole_solved"] > 0


def test_margin_flags_are_reported_not_dropped(solver):
    import bench
    w = bench.make_single(dict(bench.SINGLE["cfg2"], seed=9300), 3, capi.dt_initial)
    args = (w["N"], w["ff"], w["x0"], w["xf"], w["lim"], w["poly_ofs"], w["face_ofs"], w["Ab"], w["cand_ofs"], w["dt"], w["sigma"])
    f0, near = cr.margin_flags(solver, *args, eps=1e-2)      # a coarse band so that some candidates fall into it
    assert len(near) == len(f0) == 3 * 1024 and near.any() and not near.all()
    rep = cr.margin_report(solver, *args)
    assert rep["candidates"] == 3 * 1024 and rep["within_1e-06"] <= rep["within_1e-05"] <= rep["within_0.0001"]
