"""Generates tests/golden/mm_mid.npz + mm_mid_golden.json: a handful of MID-SIZE real Maros-Meszaros problems (n_kkt 1 500 .. 6 200)
from the reference's fixtures (/root/reference/tests/data/maros_meszaros/*.mat).  They exercise what the small subset cannot:
fronts beyond shared memory in the CTA-per-QP schedule and, for the ones whose largest front reaches 512 rows, the whole-GPU
schedule on structured (non-random) patterns.  Golden values: the oracle's status / iterations / objective with default settings,
kkt_solver = sparse_ldlt and the product's own fill-reducing ordering (the oracle's exact minimum degree is too slow here).

    python tests/golden/make_mm_mid.py        (build container only)"""
import json
import os
import sys
import time
import warnings

import numpy as np
import scipy.io
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
SRC = "/root/reference/tests/data/maros_meszaros"
NAMES = ["CVXQP1_M", "CVXQP2_M", "CVXQP3_M", "STCQP2", "CONT-050", "AUG3DCQP", "QSHIP08L", "LISWET1", "DTOC3", "STADAT1"]


def main():
    from oracle import pyoracle
    import piqp_b200
    arrays, golden = {}, []
    for name in NAMES:
        path = os.path.join(SRC, name + ".mat")
        if not os.path.exists(path):
            print("missing", name); continue
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            d = scipy.io.loadmat(path)
        g = lambda k: np.asarray(d[k], dtype=float).ravel()
        P, A, G = sp.csc_matrix(d["P"]), sp.csc_matrix(d["A"]), sp.csc_matrix(d["G"])
        n, p, m = P.shape[0], A.shape[0], G.shape[0]
        for M in (P, A, G):
            M.sort_indices()
        args = (P, g("c"), A if p else None, g("b") if p else None, G if m else None, g("h_l") if m else None, g("h_u") if m else None, g("x_l"), g("x_u"))
        t0 = time.time()
        o = pyoracle.SparseSolver(pyoracle.default_settings(kkt_solver="sparse_ldlt"), identity_preconditioner=False)
        # ordering from the product's host-only symbolic phase on the SCALED pattern (same pattern as the unscaled one)
        perm = piqp_b200.sparse_ldlt_symbolic(sp.triu(P), A if p else None, G if m else None)["perm"]
        o = pyoracle.SparseSolver(pyoracle.default_settings(kkt_solver="sparse_ldlt"), kkt_perm=perm)
        o.setup(*args)
        status = o.solve(); r = o.result()
        sym = piqp_b200.sparse_ldlt_symbolic(sp.triu(P), A if p else None, G if m else None)
        print("%-10s n=%5d p=%5d m=%5d status=%d iter=%d obj=%.10g fmax=%d  %.1fs" % (name, n, p, m, status, r.info.iter, r.info.primal_obj, sym["largest_front"], time.time() - t0), flush=True)
        if status != 1:
            continue
        golden.append({"name": name, "n": n, "p": p, "m": m, "status": int(status), "iter": int(r.info.iter), "primal_obj": float(r.info.primal_obj),
                       "largest_front": int(sym["largest_front"])})
        for key, M in (("P", P), ("A", A), ("G", G)):
            arrays["%s/%s_indptr" % (name, key)] = M.indptr.astype(np.int32)
            arrays["%s/%s_indices" % (name, key)] = M.indices.astype(np.int32)
            arrays["%s/%s_data" % (name, key)] = M.data.astype(np.float64)
        for key in ("c", "b", "h_l", "h_u", "x_l", "x_u"):
            arrays["%s/%s" % (name, key)] = g(key)
        arrays["%s/dims" % name] = np.array([n, p, m], dtype=np.int32)
    np.savez_compressed(os.path.join(HERE, "mm_mid.npz"), **arrays)
    json.dump({"source": SRC, "reference_assertion": "tests/src/sparse/maros_meszaros_tests.cpp:35 ASSERT_EQ(status, PIQP_SOLVED)", "problems": golden},
              open(os.path.join(HERE, "mm_mid_golden.json"), "w"), indent=1)
    print("%d problems, %.0f KB" % (len(golden), os.path.getsize(os.path.join(HERE, "mm_mid.npz")) / 1024))


if __name__ == "__main__":
    main()
