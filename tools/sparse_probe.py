"""Probe for the sparse workload: one batched solve (for ncu captures) + timing of setup / solve."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
class A: n=500; p=100; m=250; density=0.01
if len(sys.argv) > 1: A.n, A.p, A.m = int(sys.argv[1]), int(sys.argv[1]) // 5, int(sys.argv[1]) // 2
if len(sys.argv) > 2: A.density = float(sys.argv[2])
B = int(sys.argv[3]) if len(sys.argv) > 3 else 64
wl = bench.SparseWorkload(A)
dev = torch.device("cuda", 0)
t0 = time.perf_counter(); data = wl.device_data(B, 42, dev); torch.cuda.synchronize(); t1 = time.perf_counter()
s = wl.make_solver(0, data); torch.cuda.synchronize(); t2 = time.perf_counter()
s.set_profiling(True)
infos = s.solve(); torch.cuda.synchronize(); t3 = time.perf_counter()
st = s.stats()
print("gen %.1f ms  setup %.1f ms  solve %.1f ms  iters(max) %d  status %s" % (1e3*(t1-t0), 1e3*(t2-t1), 1e3*(t3-t2), max(i.iter for i in infos), sorted(set(i.status for i in infos))))
print("factor calls %d  factor kernel %.3f ms/call   backend solves %d  %.3f ms/call  total %.1f ms" % (st.cholesky_calls, st.cholesky_ms / max(1, st.cholesky_calls), st.backend_solves, st.backend_solve_ms / max(1, st.backend_solves), st.total_ms))
