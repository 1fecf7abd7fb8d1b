#!/usr/bin/env python
"""Kernel-only A/B of library variants on one box (run under gpurun): for each library (names under faster_b200/lib) the
two sweep launches of cfg4 batch 0 alone (CUDA events, cold L2, 20 repetitions), the flags' checksum and the iteration
histogram.  usage: kernel_ab.py lib1.so lib2.so ...   (each variant runs in its own process: FQ_LIB is read at import)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch
    import bench
    from faster_b200 import capi
    dev = torch.device("cuda:0")
    solver = capi.Solver(0)
    torch.cuda.set_stream(torch.cuda.Stream(device=dev))     # a real stream: 0 would mean "the context's own" to the library
    b = bench.PairBatch(bench.load_cfg4(0, 64), dev, torch, capi)
    solver.replan_pairs_dev(b.args, 0, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    res = b.results(capi)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = {}
    for kind, N, ff in (("whole", 10, True), ("safe", 10, False)):
        e = bench.expanded_arrays(b, res, kind, torch, dev)
        d = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in e.items()}
        nc = len(e["dt"])
        feas = torch.zeros(nc, dtype=torch.uint8, device=dev); cost = torch.zeros(nc, dtype=torch.float64, device=dev)
        its = torch.zeros(nc, dtype=torch.int32, device=dev)
        solver.set_option("max_faces_per_polytope", b.w["max_poly_faces_" + kind])
        st = torch.cuda.current_stream()

        def launch(wi):
            solver.solve_multi_dev(N, ff, b.w["n_prob"], d["x0"].data_ptr(), d["xf"].data_ptr(), b.d["lim"].data_ptr(),
                                   b.d["poly_ofs_" + kind].data_ptr(), b.d["face_ofs_" + kind].data_ptr(), b.d["Ab_" + kind].data_ptr(),
                                   d["cand_ofs"].data_ptr(), nc // b.w["n_prob"], b.w["max_faces_" + kind], d["dt"].data_ptr(),
                                   d["sigma"].data_ptr(), feas.data_ptr(), cost.data_ptr(), 0, its.data_ptr() if wi else 0, st.cuda_stream)
        torch.cuda.synchronize()
        launch(False)
        ms = []
        for _ in range(20):
            flush.zero_()
            a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(st); launch(False); z.record(st)
            torch.cuda.synchronize()
            ms.append(a.elapsed_time(z))
        launch(True)
        torch.cuda.synchronize()
        f = feas.cpu().numpy(); it = its.cpu().numpy()
        out[kind] = dict(ms_mean=float(np.mean(ms)), ms_min=float(np.min(ms)), feasible=int(f.sum()),
                         flags_crc=int(np.bitwise_xor.reduce(np.packbits(f).astype(np.uint64) * np.arange(1, len(f) // 8 + 1, dtype=np.uint64))),
                         infeasible_at_iter_1=float(((f == 0) & (it == 1)).mean()), mean_iters=float(np.abs(it).mean()))
    print(json.dumps(out))
    sys.exit(0)

for r in range(int(os.environ.get("AB_ROUNDS", "2"))):
    for lib in sys.argv[1:]:
        env = dict(os.environ, FQ_LIB=os.path.join(ROOT, "faster_b200", "lib", lib))
        p = subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env=env, capture_output=True, text=True)
        try:
            d = json.loads(p.stdout.strip().splitlines()[-1])
            print("%-28s whole %.4f ms (min %.4f)  safe %.4f ms (min %.4f)  refuted at iteration 1: %.3f / %.3f  flags %x/%x" %
                  (lib, d["whole"]["ms_mean"], d["whole"]["ms_min"], d["safe"]["ms_mean"], d["safe"]["ms_min"],
                   d["whole"]["infeasible_at_iter_1"], d["safe"]["infeasible_at_iter_1"], d["whole"]["flags_crc"], d["safe"]["flags_crc"]), flush=True)
        except Exception as ex:
            print(lib, "FAILED", ex, p.stderr[-600:], flush=True)
