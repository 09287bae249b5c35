"""Generates tests/golden/mm_small.npz + mm_small_golden.json from the reference's own Maros-Meszaros fixtures
(/root/reference/tests/data/maros_meszaros/*.mat, the files tests/src/sparse/maros_meszaros_tests.cpp:20-36 iterates over and
asserts PIQP_SOLVED on with default settings).  Run in the build container (the reference tree is not on the GPU box):

    python tests/golden/make_mm_small.py

Keeps the problems with n + p + m <= 450 (small enough to commit; 33 of the 138).  For each it stores the problem data
(CSC arrays, vectors with +-inf bounds as in the file) and, in the json, what the CPU oracle computed for it with default
settings and kkt_solver = sparse_ldlt: status, iterations, primal objective.  The status column is pinned by the reference
(its test asserts PIQP_SOLVED = 1 for every file); iterations / objective are the oracle's and serve as regression values
for the CUDA path (same status, same iteration count, |dx| <= 1e-8 max(1, |x|))."""
import glob
import json
import os
import sys
import warnings

import numpy as np
import scipy.io
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
SRC = "/root/reference/tests/data/maros_meszaros"
LIMIT = 450


def load(path):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        d = scipy.io.loadmat(path)
    g = lambda k: np.asarray(d[k], dtype=float).ravel()
    return sp.csc_matrix(d["P"]), g("c"), sp.csc_matrix(d["A"]), g("b"), sp.csc_matrix(d["G"]), g("h_l"), g("h_u"), g("x_l"), g("x_u")


def main():
    from oracle import pyoracle
    arrays, golden = {}, []
    for path in sorted(glob.glob(os.path.join(SRC, "*.mat"))):
        name = os.path.basename(path)[:-4]
        P, c, A, b, G, h_l, h_u, x_l, x_u = load(path)
        n, p, m = P.shape[0], A.shape[0], G.shape[0]
        if n + p + m > LIMIT:
            continue
        for M in (P, A, G):
            M.sort_indices()
        s = pyoracle.SparseSolver(pyoracle.default_settings(kkt_solver="sparse_ldlt"))
        s.setup(P, c, A if p else None, b if p else None, G if m else None, h_l if m else None, h_u if m else None, x_l, x_u)
        status = s.solve()
        r = s.result()
        golden.append({"name": name, "n": n, "p": p, "m": m, "status": int(status), "iter": int(r.info.iter), "primal_obj": float(r.info.primal_obj)})
        for key, M in (("P", P), ("A", A), ("G", G)):
            arrays["%s/%s_indptr" % (name, key)] = M.indptr.astype(np.int32)
            arrays["%s/%s_indices" % (name, key)] = M.indices.astype(np.int32)
            arrays["%s/%s_data" % (name, key)] = M.data.astype(np.float64)
        for key, v in (("c", c), ("b", b), ("h_l", h_l), ("h_u", h_u), ("x_l", x_l), ("x_u", x_u)):
            arrays["%s/%s" % (name, key)] = v
        arrays["%s/dims" % name] = np.array([n, p, m], dtype=np.int32)
        print("%-10s n=%4d p=%4d m=%4d status=%d iter=%d obj=%.10g" % (name, n, p, m, status, r.info.iter, r.info.primal_obj))
    np.savez_compressed(os.path.join(HERE, "mm_small.npz"), **arrays)
    json.dump({"source": SRC, "limit_n_p_m": LIMIT, "reference_assertion": "tests/src/sparse/maros_meszaros_tests.cpp:35 ASSERT_EQ(status, PIQP_SOLVED)",
               "problems": golden}, open(os.path.join(HERE, "mm_small_golden.json"), "w"), indent=1)
    print("%d problems, %.0f KB" % (len(golden), os.path.getsize(os.path.join(HERE, "mm_small.npz")) / 1024))


if __name__ == "__main__":
    main()
