"""Iteration-count parity of the dense CUDA path against the CPU oracle at BASELINE config 2's shape (n=1024, m=512):
B instances (the seeds of tests/test_gpu_dense.py: 42 + b) solved as one batch on the GPU and one by one by the oracle (16 host threads).
    python tools/dense_iter_parity.py [B]        prints per instance: iterations GPU / oracle, |dx|_inf / max(1, |x|_inf)"""
import os
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    import piqp_b200
    from piqp_b200.synth import dense_strongly_convex_qp
    from helpers import setup_args
    from oracle import pyoracle
    pyoracle.lib()          # the portable build the parity tests use (no -march=native: FMA contraction would change the oracle's rounding)
    qs = [dense_strongly_convex_qp(1024, 0, 512, seed=42 + b) for b in range(B)]
    s = piqp_b200.DenseSolverBatched()
    stack = lambda k: None if qs[0].get(k) is None else np.stack([q[k] for q in qs])
    s.setup(*[stack(k) for k in ("P", "c", "A", "b", "G", "h_l", "h_u", "x_l", "x_u")])
    s.solve()
    r = s.result()

    def one(q):
        o = pyoracle.DenseSolver(); o.setup(*setup_args(q)); st = o.solve(); ro = o.result()
        return st, ro.info.iter, ro.x
    with ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 1)) as ex:
        res = list(ex.map(one, qs))
    bad = 0
    for b, (st, it, x) in enumerate(res):
        dx = np.abs(r.x[b] - x).max() / max(1.0, np.abs(x).max())
        ok = r.info[b].iter == it and r.info[b].status == st and dx <= 1e-8
        bad += not ok
        print("instance %3d: status %d / %d  iterations %3d / %3d  dx %.2e %s" % (b, r.info[b].status, st, r.info[b].iter, it, dx, "" if ok else "  <-- MISMATCH"))
    print("%d / %d instances match the oracle in status, iteration count and 1e-8" % (B - bad, B))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
