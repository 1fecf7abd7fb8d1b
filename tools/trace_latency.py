import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
import bench
from faster_b200 import capi, corridor as cr
s = capi.Solver(0)
w = bench.load_cfg4(0, 1)
polys = [(w["Ab_whole"][w["face_ofs_whole"][p]:w["face_ofs_whole"][p + 1], :3], w["Ab_whole"][w["face_ofs_whole"][p]:w["face_ofs_whole"][p + 1], 3]) for p in range(3)]
x0, xf, lim = w["x0"][0], w["xf_whole"][0], w["lim"][0]
sig66 = cr.monotone_sigmas(10, 3)
dts10 = np.arange(1.0, 11.0) * max(capi.dt_initial(x0, xf, lim, 10), 0.02)
for ee in (0, 1):
    s.set_option("sweep_early_exit", ee)
    print("early exit", ee, file=sys.stderr, flush=True)
    for i in range(12):
        g = s.gen_new_traj(10, x0, xf, lim, polys, dts10, sig66, True)
print(g["dt_index"], g["cost"])
