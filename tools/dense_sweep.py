"""Dense size sweep (north star: "synthetic dense n in {128 ... 4096}"): batched device-resident solves, both assembly kernels.
Prints one line per (n, mode): QP/s, algorithmic factor+solve TFLOP/s, assembly ms per launch, Cholesky ms per call."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

out = []
sizes = [int(a) for a in sys.argv[1:]] or [128, 256, 512, 1024, 2048, 4096]
for n in sizes:
    class A: pass
    A.n, A.p, A.m = n, 0, n // 2
    B = max(8, min(1024, int(256 * (1024 / n) ** 2)))
    for mode in ("dmma", "ozaki"):
        os.environ["B200_DENSE_ASSEMBLE"] = mode
        wl = bench.DenseWorkload(A)
        dev = torch.device("cuda", 0)
        data = wl.device_data(B, 42, dev)
        s = wl.make_solver(0, data)
        s.set_profiling(True)
        s.solve()                                   # warm-up
        infos = s.solve(); torch.cuda.synchronize()
        st = s.stats()
        ff, sf = wl.work()
        fl = st.factor_calls * ff + st.backend_solves * sf
        row = dict(n=n, m=n // 2, batch=B, mode=mode, qps=B / (st.total_ms * 1e-3), tflops=fl / (st.total_ms * 1e-3) * 1e-12, step_ms=st.total_ms,
                   assemble_ms=st.assemble_ms / max(1, st.assemble_launches), cholesky_ms=st.cholesky_ms / max(1, st.cholesky_calls),
                   iters=max(i.iter for i in infos), solved=all(i.status == 1 for i in infos))
        out.append(row)
        print(json.dumps(row), flush=True)
        del s, data
        torch.cuda.empty_cache()
