"""CPU-side checks of the chained-replan / multi-GPU part of the ABI: struct layouts, the shard rule, the device/host
shared arithmetic header, and the multi-process exchange logic with gloo (world size 2)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

from faster_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_struct_layouts_match_the_header(built_lib, tmp_path):
    """sizeof/offsetof of fq_pair_args and fq_pair_result as the C compiler sees them vs the ctypes mirror."""
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "faster_b200.h"\n'
                   'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(fq_pair_result), sizeof(fq_pair_args),'
                   'offsetof(fq_pair_result, R), offsetof(fq_pair_result, k_safe), offsetof(fq_pair_args, results),'
                   'offsetof(fq_pair_args, sigmas_safe), offsetof(fq_pair_args, max_poly_faces_safe)); return 0;}\n')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    want = [C.sizeof(capi.PairResult), C.sizeof(capi.PairArgs), capi.PairResult.R.offset, capi.PairResult.k_safe.offset,
            capi.PairArgs.results.offset, capi.PairArgs.sigmas_safe.offset, capi.PairArgs.max_poly_faces_safe.offset]
    assert got == want
    assert capi.PAIR_RESULT_DTYPE.fields["R"][1] == capi.PairResult.R.offset


def test_shard_range_partitions_every_problem_once():
    rng = np.random.default_rng(0)
    for _ in range(200):
        n = int(rng.integers(1, 40)); world = int(rng.integers(1, 9))
        co = np.concatenate([[0], np.cumsum(rng.integers(0, 2000, n))]).astype(np.int32)
        for ofs in (None, co):
            cuts = [capi.shard_range(n, ofs, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            for a, b in zip(cuts[:-1], cuts[1:]):
                assert a[1] == b[0] and a[0] <= a[1]
        # balanced by candidates: no shard exceeds its fair share by more than the largest problem
        if co[-1] > 0:
            big = int(np.diff(co).max())
            for r in range(world):
                lo, hi = capi.shard_range(n, co, r, world)
                assert co[hi] - co[lo] <= co[-1] / world + big
    with pytest.raises(capi.FqError):
        capi.shard_range(4, None, 3, 2)


def test_shared_arithmetic_header_host_side(oracle):
    """fq_dtinit.h compiled for the host (fq_dt_initial, fq_fill_x, fq_num_samples) equals the oracle's restatement
    bit for bit on random states (the device compilation of the same source is compared on the GPU)."""
    rng = np.random.default_rng(3)
    for _ in range(2000):
        x0 = np.concatenate([rng.uniform(-5, 5, 3), rng.uniform(-4, 4, 3), rng.uniform(-3, 3, 3)])
        xf = np.concatenate([x0[:3] + rng.uniform(-5, 5, 3), np.zeros(6)])
        if rng.random() < 0.1:
            x0[3:] = 0
        lim = np.array([5.0, 5.0, 8.0]) if rng.random() < 0.7 else np.array([1.4, 1.4, 5.0])
        N = int(rng.integers(3, 16))
        assert capi.dt_initial(x0, xf, lim, N) == oracle.dt_initial(x0, xf, lim, N)
    co = rng.normal(size=(10, 12))
    for dt in (0.05, 0.31, 0.5003):
        assert capi.num_samples(10, dt, 0.01) == oracle.lib().fqo_num_samples(10, dt, 0.01)
        assert np.array_equal(capi.fill_x(10, co, dt, 0.01), oracle.fill_x(10, co, dt, 0.01))


def test_multi_gpu_entry_points_fail_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.FqError):
        capi.Solver(n_gpus=2)
    assert len(capi.comm_unique_id()) == 128 or True      # NCCL may or may not load on a CPU-only box; must not crash


def _gloo_worker(rank, world, port, q):
    """What bench.py / a launcher does around the library's collective, with gloo standing in for NCCL: every rank solves
    its shard (here: fabricates its records), all-gathers the fixed-size result records, and every rank ends up with the
    same full table in rank order."""
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 10
    lo, hi = capi.shard_range(n, None, rank, world)
    pad = max(capi.shard_range(n, None, r, world)[1] - capi.shard_range(n, None, r, world)[0] for r in range(world))
    rec = np.zeros(pad, capi.PAIR_RESULT_DTYPE)
    rec["whole_dt_index"][:hi - lo] = np.arange(lo, hi)
    rec["whole_cost"][:hi - lo] = 100.0 + np.arange(lo, hi)
    send = torch.from_numpy(rec.view(np.uint8).copy())
    recv = [torch.zeros_like(send) for _ in range(world)]
    dist.all_gather(recv, send)
    table = np.zeros(n, capi.PAIR_RESULT_DTYPE)
    for r in range(world):
        a, b = capi.shard_range(n, None, r, world)
        table[a:b] = recv[r].numpy().view(capi.PAIR_RESULT_DTYPE)[:b - a]
    q.put((rank, table.tobytes()))
    dist.destroy_process_group()


def test_result_record_exchange_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    ps = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in ps:
        p.join(timeout=60)
    assert got[0] == got[1]
    t = np.frombuffer(got[0], capi.PAIR_RESULT_DTYPE)
    assert np.array_equal(t["whole_dt_index"], np.arange(10)) and np.array_equal(t["whole_cost"], 100.0 + np.arange(10))
