"""Probe for the dense workload: one batched solve of config 2 (for ncu captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
class A: n=1024; p=0; m=512
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
wl = bench.DenseWorkload(A)
dev = torch.device("cuda", 0)
data = wl.device_data(B, 42, dev)
s = wl.make_solver(0, data)
s.set_profiling(True)
infos = s.solve(); torch.cuda.synchronize()
st = s.stats()
print("iters(max) %d status %s  assemble %.3f ms/launch  cholesky %.3f ms/call  total %.1f ms" % (
    max(i.iter for i in infos), sorted(set(i.status for i in infos)), st.assemble_ms / max(1, st.assemble_launches), st.cholesky_ms / max(1, st.cholesky_calls), st.total_ms))
