"""Timing probe for the multistage workload: fresh-solver setup vs solve (host wall clock), iterations."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import piqp_b200
from piqp_b200.synth import mpc_batch

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
d = mpc_batch(B)
for rep in range(reps):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    s = piqp_b200.SparseSolverBatched(device=0, kkt_solver="sparse_multistage")
    s.setup(B, d["P"], d["c"], d["A"], d["b"], None, None, None, d["x_l"], d["x_u"], Ax=d["Ax"])
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    infos = s.solve()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    infos2 = s.solve()
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    del s
    torch.cuda.synchronize()
    t4 = time.perf_counter()
    print("rep %d: setup %.1f ms  solve#1 %.1f ms  solve#2 %.1f ms  del %.1f ms  iters %s status %s" % (
        rep, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), 1e3 * (t4 - t3), sorted(set(i.iter for i in infos)), sorted(set(i.status for i in infos))))
