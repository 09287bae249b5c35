"""Timing of the dense LDL^T without pivoting (benchmarks/src/dense_cholesky_factorization_benchmark.cpp:76-95 times the
reference's LDLTNoPivot) through piqp_b200.LDLTNoPivot: usage  python tools/dense_ldlt_probe.py [n ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import piqp_b200

for n in [int(a) for a in sys.argv[1:]] or [1024, 2048, 4096]:
    rng = np.random.default_rng(n)
    M = rng.standard_normal((n, n)); P = M @ M.T / n + np.eye(n)
    L = np.tril(P)
    ldlt = piqp_b200.LDLTNoPivot()
    t0 = time.perf_counter(); ldlt.compute(L); t1 = time.perf_counter()
    kkt = ldlt._kkt
    x_reg = np.zeros(n); z0 = np.zeros(0)
    ts = []
    for _ in range(3):
        t2 = time.perf_counter(); ok = kkt.update_scalings_and_factor(1.0, x_reg, z0); ts.append(time.perf_counter() - t2)
    b = rng.standard_normal(n)
    t4 = time.perf_counter(); x = ldlt.solve(b); t5 = time.perf_counter()
    res = np.abs(P @ x - b).max()
    fl = n ** 3 / 3.0
    print("n=%5d  create+first factor %.2f s   factor %.2f ms (%.2f TFLOP/s, incl. the C-ABI call)   solve %.2f ms   residual %.1e  ok=%s"
          % (n, t1 - t0, 1e3 * min(ts), fl / min(ts) * 1e-12, 1e3 * (t5 - t4), res, ok), flush=True)
