import sys, time, numpy as np
sys.path.insert(0,'.')
from faster_b200 import capi, corridor as cr
s=capi.Solver(0)
pb=cr.make_corridor(10000,3,10)
sig=cr.monotone_sigmas(10,3)
dti=capi.dt_initial(pb["x0"],pb["xf"],pb["lim"],10)
dts=np.arange(1.0,11.0)*max(dti,0.02)
for _ in range(20): g=s.gen_new_traj(10,pb["x0"],pb["xf"],pb["lim"],pb["polys"],dts,sig,True)
lat=[]
for _ in range(100):
    t=time.perf_counter(); g=s.gen_new_traj(10,pb["x0"],pb["xf"],pb["lim"],pb["polys"],dts,sig,True); lat.append(time.perf_counter()-t)
print('gen_new_traj median us', np.median(lat)*1e6, 'min', np.min(lat)*1e6, g['dt_index'], g['cost'])
fe,co,_,it=s.solve_batch(10,pb["x0"],pb["xf"],pb["lim"],pb["polys"],np.repeat(dts,len(sig)),np.tile(sig,(10,1)),True,False,True)
print('iters max/mean', it.max(), it.mean(), 'feasible', fe.mean())
