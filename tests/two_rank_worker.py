"""Worker of tests/test_pair_gpu.py::test_two_ranks_one_process_each: rank r owns GPU r, attaches its context to the
communicator (the id travels through a file, as a launcher would hand it over) and runs the chained replan with the
all-gather of the result records."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank, world, uid_path, out_path = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4]
    import torch
    from faster_b200 import capi
    from test_pair_gpu import _synthetic_pairs
    if rank == 0:
        uid = capi.comm_unique_id()
        with open(uid_path + ".tmp", "wb") as f:
            f.write(uid)
        os.rename(uid_path + ".tmp", uid_path)
    else:
        for _ in range(600):
            if os.path.exists(uid_path):
                break
            time.sleep(0.1)
        uid = open(uid_path, "rb").read()
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    s = capi.Solver(rank)
    s.comm_init(uid, rank, world)
    assert s.comm_info()[:2] == (rank, world)
    w = _synthetic_pairs(5, 7700 + 1000 * rank)
    d = {k: torch.from_numpy(np.ascontiguousarray(w[k])).to(dev) for k in capi.PAIR_INPUT_KEYS}
    n = w["n_prob"]
    res = torch.zeros(n * 144, dtype=torch.uint8, device=dev)
    allres = torch.zeros(world * n * 144, dtype=torch.uint8, device=dev)
    a = capi.pair_args(w, lambda k: d[k].data_ptr(), {"results": res.data_ptr()})
    torch.cuda.synchronize()        # the fills above run on torch's stream, the library on the context's
    s.replan_pairs_dev(a, allres.data_ptr(), 0)
    s.wait()
    np.save(out_path, np.frombuffer(allres.cpu().numpy().tobytes(), capi.PAIR_RESULT_DTYPE))
    s.close()


if __name__ == "__main__":
    main()
