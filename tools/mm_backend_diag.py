"""diagnostic: accuracy of the CUDA sparse backends on the KKT systems of the chaotic Maros-Meszaros problems at interior-point-like
scalings (delta = 1e-10, z_reg spread over 16 decades): residual of the full 3x3 system for the multifrontal kernels, the
level-scheduled kernels and the oracle (same permutation)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import scipy.sparse as sp
import piqp_b200
from oracle import pyoracle
from helpers import load_mm_small
P, G = load_mm_small()
for name in sys.argv[1:] or ["QBEACONF", "QRECIPE", "QSC205"]:
    o = pyoracle.SparseSolver(pyoracle.default_settings(kkt_solver="sparse_ldlt")); o.setup(*P[name])
    Ps, AT, GT = o.scaled_matrices(); n, p, m = o.dims[:3]
    rng = np.random.default_rng(0)
    for delta, lo, hi in ((1e-4, -2, 2), (1e-10, -8, 8)):
        x_reg = np.full(n, delta) + 10.0 ** rng.uniform(lo, hi, n) * (rng.random(n) < 0.5)
        z_reg = 10.0 ** rng.uniform(lo, hi, m)
        Pf = sp.csc_matrix(Ps) + sp.triu(sp.csc_matrix(Ps), 1).T
        K = sp.bmat([[Pf + sp.diags(x_reg), AT, GT], [AT.T, -delta * sp.eye(p), None], [GT.T, None, -sp.diags(z_reg) if m else None]]).tocsc()
        rhs = (rng.standard_normal(n), rng.standard_normal(p), rng.standard_normal(m))
        r = np.concatenate(rhs)
        out = []
        perm = None
        for label, env in (("frontal", {"B200_LDLT_LEVELS": "0"}), ("levels", {"B200_LDLT_LEVELS": "1"}), ("frontal_noamalg", {"B200_LDLT_LEVELS": "0", "B200_LDLT_NO_AMALG": "1"})):
            for k2, v2 in env.items():
                os.environ[k2] = v2
            be = piqp_b200.SparseKKT(Ps, AT, GT)
            ok = be.update_scalings_and_factor(delta, x_reg, z_reg)
            sol = np.concatenate(be.solve(*rhs))
            out.append((label, ok, "%.2e" % (np.abs(K @ sol - r).max() / max(1.0, np.abs(sol).max())), "%.3e" % np.abs(sol).max()))
            if label == "frontal":
                perm = be.symbolic_info()["perm"]; sol_f = sol
            if label == "levels":
                out[-1] += ("vs frontal %.2e" % (np.abs(sol - sol_f).max() / max(1.0, np.abs(sol).max())),)
            os.environ.pop("B200_LDLT_NO_AMALG", None)
        o2 = pyoracle.SparseSolver(pyoracle.default_settings(kkt_solver="sparse_ldlt"), kkt_perm=perm); o2.setup(*P[name])
        o2.backend_factor(delta, x_reg, z_reg)
        so = np.concatenate(o2.backend_solve(*rhs))
        out.append(("oracle", "%.2e" % (np.abs(K @ so - r).max() / max(1.0, np.abs(so).max())), "vs frontal %.2e" % (np.abs(so - sol_f).max() / max(1.0, np.abs(so).max()))))
        print(name, "delta=%g" % delta, out, flush=True)
