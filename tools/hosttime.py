import os, sys, time, numpy as np
sys.path.insert(0,'.')
import bench
from faster_b200 import capi
import torch
s=capi.Solver(0)
w=bench.make_workload(64,10000,'whole')
keys=["x0","xf","lim","poly_ofs","face_ofs","Ab","cand_ofs","dt","sigma"]
h={k: torch.from_numpy(np.ascontiguousarray(w[k])).pin_memory() for k in keys}
n=64*1024
o=(torch.zeros(n,dtype=torch.uint8).pin_memory(), torch.zeros(n,dtype=torch.float64).pin_memory())
def call():
    s.solve_multi(10, True, h["x0"].numpy(), h["xf"].numpy(), h["lim"].numpy(), h["poly_ofs"].numpy(), h["face_ofs"].numpy(), h["Ab"].numpy(), h["cand_ofs"].numpy(), h["dt"].numpy(), h["sigma"].numpy(), out=(o[0].numpy(),o[1].numpy(),None,None))
for _ in range(5): call()
t=time.perf_counter()
for _ in range(100): call()
print('e2e whole call ms', (time.perf_counter()-t)*10)
