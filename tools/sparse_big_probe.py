"""Probe of the sparse_ldlt backend on ONE large random sparse QP (BASELINE config 3 family: n, p = m = n/2, density d):
symbolic time, factor / solve time through the C-ABI (b200kkt_*), residual of the solve against the KKT matrix.
usage: python tools/sparse_big_probe.py [n=10000] [density=0.01] [mode=0] [reps=2]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scipy.sparse as sp
import piqp_b200
from piqp_b200.synth import sparse_strongly_convex_qp

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
dens = float(sys.argv[2]) if len(sys.argv) > 2 else 0.01
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 0
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
p = m = n // 2
q = sparse_strongly_convex_qp(n, p, m, dens, seed=42)
P = sp.csc_matrix(sp.triu(q["P"])); AT = sp.csc_matrix(q["A"].T); GT = sp.csc_matrix(q["G"].T)
t0 = time.perf_counter()
be = piqp_b200.SparseKKT(P, AT, GT, mode=mode)
t1 = time.perf_counter()
info = be.symbolic_info()
print("n=%d p=%d m=%d density=%g mode=%d: create (symbolic + upload) %.2f s, nnz(KKT)=%d nnz(L)=%d" % (n, p, m, dens, mode, t1 - t0, info["nnz_kkt"], info["nnz_L"]), flush=True)
be.print_info()
rng = np.random.default_rng(0)
x_reg = rng.uniform(0.5, 1.5, n); z_reg = rng.uniform(0.5, 2.0, m); delta = 0.9
for r in range(reps):
    t0 = time.perf_counter(); ok = be.update_scalings_and_factor(delta, x_reg, z_reg); t1 = time.perf_counter()
    rhs = (rng.standard_normal(n), rng.standard_normal(p), rng.standard_normal(m))
    t2 = time.perf_counter(); sol = be.solve(*rhs); t3 = time.perf_counter()
    print("rep %d: factor ok=%s %.1f ms   solve %.1f ms" % (r, ok, 1e3 * (t1 - t0), 1e3 * (t3 - t2)), flush=True)
Pf = P + sp.triu(P, 1).T
K = sp.bmat([[Pf + sp.diags(x_reg), AT, GT], [AT.T, -delta * sp.eye(p), None], [GT.T, None, -sp.diags(z_reg)]]).tocsc()
s = np.concatenate(sol)
res = np.abs(K @ s - np.concatenate(rhs)).max()
print("residual |K sol - rhs|_inf = %.3e  (|sol|_inf = %.3e)" % (res, np.abs(s).max()))
assert res <= 1e-8 * max(1.0, np.abs(s).max())
print("ok")
