import sys, time, numpy as np
sys.path.insert(0,'.')
from faster_b200 import capi, corridor as cr
s=capi.Solver(0)
pb=cr.make_corridor(10000,3,10)
sig=cr.monotone_sigmas(10,3)
dti=capi.dt_initial(pb["x0"],pb["xf"],pb["lim"],10)
dts=np.arange(1.0,11.0)*max(dti,0.02)
def timeit(f, n=200):
    for _ in range(20): f()
    lat=[]
    for _ in range(n):
        t=time.perf_counter(); f(); lat.append(time.perf_counter()-t)
    return np.median(lat)*1e6
def host():
    g=s.gen_new_traj(10,pb["x0"],pb["xf"],pb["lim"],pb["polys"],dts,sig,True)
    return capi.fill_x(10,g["coeffs"],dts[g["dt_index"]],0.01)
def dev():
    return s.gen_new_traj_sampled(10,pb["x0"],pb["xf"],pb["lim"],pb["polys"],dts,sig,0.01,True,max_samples=1024)["X"]
print('genNewTraj+fillX: host sampling %.1f us, device sampling %.1f us, samples %d' % (timeit(host), timeit(dev), len(host())))
def exact():
    return s.gen_new_traj_exact(10,pb["x0"],pb["xf"],pb["lim"],pb["polys"],dts,True)
g=exact()
print('exact sweep: %.1f us, nodes %d, exact %s, dt_index %d' % (timeit(exact, 100), g['nodes'], g['exact'], g['dt_index']))
# a loose time allocation (factor 5): bigger tree
dts5=np.array([5.0*max(dti,0.02)])
def exact5():
    return s.gen_new_traj_exact(10,pb["x0"],pb["xf"],pb["lim"],pb["polys"],dts5,True)
g=exact5()
print('exact single dt (factor 5): %.1f us, nodes %d' % (timeit(exact5, 50), g['nodes']))
