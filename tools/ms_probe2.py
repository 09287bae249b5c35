"""e2e-style probe (as bench.py does it): pinned host tensors, fresh solver per repetition, other solver alive"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ctypes as C
import piqp_b200
import bench
class A: horizon=100; nx=12; nu=4
wl = bench.MultistageWorkload(A)
dev = torch.device("cuda", 0)
data = wl.device_data(128, 42, dev)
solver = wl.make_solver(0, data)
solver.solve()
host = wl.host_data(data)
for rep in range(4):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    s2 = wl.make_solver(0, host, on_host=True)
    t1 = time.perf_counter()
    s2.solve()
    t2 = time.perf_counter()
    del s2
    t3 = time.perf_counter()
    print("rep %d make_solver %.1f ms solve %.1f ms del %.1f ms" % (rep, 1e3*(t1-t0), 1e3*(t2-t1), 1e3*(t3-t2)), file=sys.stderr)
