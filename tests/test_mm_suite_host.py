"""CPU-only: the host side of tools/mm_suite.py (BASELINE config 5): the shape table parsed from BASELINE.md, the Maros-Meszaros-
shaped generator (shared pattern, R value sets) and the LPT assignment over ranks; the oracle solves the generated QPs."""
import os
import sys

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def test_shape_table_and_generator(oracle):
    import mm_suite
    from piqp_b200.distributed import lpt_assign
    shapes = mm_suite.mm_shapes()
    assert len(shapes) == 138 and ("TAME", 2, 1, 0, 4, 2, 0) in shapes
    small = [s for s in shapes if s[1] + s[2] + s[3] <= 250][:8]
    assert len(small) == 8
    for idx, (name, n, p, m, nP, nA, nG) in enumerate(small):
        P, A, G, v = mm_suite.shaped_qp(n, p, m, nP, nA, nG, 42 + idx, 2)
        assert P.shape == (n, n) and A.shape == (p, n) and G.shape == (m, n) and sp.triu(P).nnz == P.nnz
        assert v["Px"].shape == (2, P.nnz) and v["Ax"].shape == (2, A.nnz) and v["Gx"].shape == (2, G.nnz)
        assert not np.array_equal(v["Ax"][0], v["Ax"][1]) or A.nnz == 0
        mk = lambda M, x: sp.csc_matrix((x, M.indices, M.indptr), shape=M.shape)
        for k in range(2):
            s = oracle.SparseSolver(oracle.default_settings(kkt_solver="sparse_ldlt"))
            s.setup(mk(P, v["Px"][k]), v["c"][k], mk(A, v["Ax"][k]) if p else None, v["b"][k] if p else None, mk(G, v["Gx"][k]) if m else None,
                    v["h_l"][k] if m else None, v["h_u"][k] if m else None, v["x_l"][k], v["x_u"][k])
            assert s.solve() == 1, (name, k)
    owner = lpt_assign([s[1] ** 2 for s in shapes], 8)
    load = np.bincount(owner, weights=[s[1] ** 2 for s in shapes], minlength=8)
    assert len(set(owner)) == 8 and load.max() <= max(max(s[1] ** 2 for s in shapes), 4 / 3 * load.sum() / 8)
