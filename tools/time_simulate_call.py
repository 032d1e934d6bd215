import sys, time, os
sys.path.insert(0,'sde-sim-rs_b200'); sys.path.insert(0,'tests')
import torch, sde_sim_rs as S
from conftest import GBM_EQ, HESTON_EQ, grid
torch.cuda.init(); torch.zeros(1,device='cuda')
for name,args,kw in (("C1",(GBM_EQ, grid(252), 10000, {"X1":1.0}, "pseudo","euler"),{}),
                     ("C2small",(GBM_EQ, grid(252), 10000, {"X1":1.0}, "sobol","euler"),dict(scramble="xor",icdf="fast",arithmetic="fast")),
                     ("C3small",(HESTON_EQ, grid(1000), 10000, {"S":100.0,"v":0.04}, "sobol","runge-kutta"),dict(scramble="xor",icdf="fast",arithmetic="fast"))):
    for i in range(3):
        t0=time.perf_counter(); r=S.simulate(*args, seed=1, **kw); v=r.values; torch.cuda.synchronize(); t1=time.perf_counter()
        print(name, "simulate() call %d: %.2f ms"%(i,(t1-t0)*1e3), file=sys.stderr)
