"""Golden vector for BASELINE config 3 at FULL size (n = 10 000, p = m = 5 000, 1 %, n_kkt = 20 000): one factorisation + one
backend solve of the CPU oracle (oracle/oracle_sparse.hpp: restated sparse::KKT<FULL> + sparse::LDLt, no FMA contraction)
under the PRODUCT's permutation.  The oracle needs minutes per factorisation at this size (333 GFLOP scalar), too long
for a test on the GPU box, so its answer is committed:

    python tests/golden/make_c3_factor_solve.py        ->  tests/golden/c3_factor_solve.npz   (~0.5 MB)

The QP is bench.py's `sparse_c3` workload (piqp_b200.synth.sparse_batch(1, 10000, 5000, 5000, 0.01, seed0=42), instance 0);
delta, x_reg, z_reg and the right-hand side come from numpy's default_rng(20261017).  The file stores the solution, the
permutation and checksums of the scaled matrices so that tests/test_gpu_full_size.py can verify it regenerated the same
inputs before comparing (bar: 1e-10 relative, SURVEY 8c / VERDICT r01 item 1d)."""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def inputs(n=10000, p=5000, m=5000, density=0.01):
    import scipy.sparse as sp
    from piqp_b200.synth import sparse_batch
    d = sparse_batch(1, n, p, m, density, seed0=42)
    mk = lambda pat, v: sp.csc_matrix((v, pat.indices, pat.indptr), shape=pat.shape)
    args = (mk(d["P"], d["Px"][0]), d["c"][0], mk(d["A"], d["Ax"][0]), d["b"][0], mk(d["G"], d["Gx"][0]), d["h_l"][0], d["h_u"][0], d["x_l"][0], d["x_u"][0])
    rng = np.random.default_rng(20261017)
    scal = dict(delta=0.9, x_reg=rng.uniform(0.5, 1.5, n), z_reg=rng.uniform(0.5, 2.0, m),
                rx=rng.standard_normal(n), ry=rng.standard_normal(p), rz=rng.standard_normal(m))
    return args, scal


def checksum(M):
    return float(np.dot(M.data, np.cos(np.arange(M.nnz) * 0.37))), int(M.nnz)


def main():
    import piqp_b200
    from oracle import pyoracle
    args, s = inputs()
    t0 = time.time()
    perm = piqp_b200.sparse_ldlt_symbolic(args[0], args[2], args[4])["perm"]
    print("symbolic %.1f s" % (time.time() - t0), flush=True)
    o = pyoracle.SparseSolver(pyoracle.default_settings(kkt_solver="sparse_ldlt"), kkt_perm=perm)
    o.setup(*args)
    P, AT, GT = o.scaled_matrices()
    print("oracle setup %.1f s" % (time.time() - t0), flush=True)
    ok = o.backend_factor(s["delta"], s["x_reg"], s["z_reg"])
    print("oracle factor %.1f s ok=%d" % (time.time() - t0, ok), flush=True)
    lx, ly, lz = o.backend_solve(s["rx"], s["ry"], s["rz"])
    np.savez_compressed(os.path.join(HERE, "c3_factor_solve.npz"), lx=lx, ly=ly, lz=lz, perm=perm.astype(np.int32), ok=ok,
                        chk=np.array([checksum(P)[0], checksum(AT)[0], checksum(GT)[0]]), nnz=np.array([P.nnz, AT.nnz, GT.nnz]))
    print("done %.1f s" % (time.time() - t0))


if __name__ == "__main__":
    main()
