"""Host-side symbolic phase of the sparse_ldlt backend (no GPU): ordering validity / quality and nnz(L) against the
oracle's restatement of LDLt::factorize_symbolic (include/piqp/sparse/ldlt.hpp:42-99) under the SAME permutation."""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp

from helpers import load_scenario_mpc, setup_args
from piqp_b200 import _lib
from piqp_b200.synth import mpc_batch, sparse_strongly_convex_qp


def _symbolic(P, AT, GT, perm=None):
    L = _lib.lib()
    ip = C.POINTER(C.c_int)

    def csc(M, upper=False):
        M = sp.csc_matrix(M)
        if upper:
            M = sp.triu(M, format="csc")
        M.sort_indices()
        return np.ascontiguousarray(M.indptr, dtype=np.int32), np.ascontiguousarray(M.indices, dtype=np.int32)
    n, p, m = P.shape[0], AT.shape[1], GT.shape[1]
    Pp, Pi = csc(P, True); Ap, Ai = csc(AT); Gp, Gi = csc(GT)
    out = np.zeros(n + p + m, dtype=np.int32)
    nk, nl, lv, fl = C.c_longlong(), C.c_longlong(), C.c_int(), C.c_double()
    pin = None if perm is None else np.ascontiguousarray(perm, dtype=np.int32)
    rc = L.b200_sparse_ldlt_symbolic(n, p, m, Pp.ctypes.data_as(ip), Pi.ctypes.data_as(ip), Ap.ctypes.data_as(ip), Ai.ctypes.data_as(ip),
                                     Gp.ctypes.data_as(ip), Gi.ctypes.data_as(ip), None if pin is None else pin.ctypes.data_as(ip),
                                     out.ctypes.data_as(ip), C.byref(nk), C.byref(nl), C.byref(lv), C.byref(fl))
    _lib.check(rc, "b200_sparse_ldlt_symbolic")
    return {"perm": out, "nnz_kkt": nk.value, "nnz_L": nl.value, "levels": lv.value, "flops": fl.value}


def _cases():
    q, _ = load_scenario_mpc()
    yield "notebook", setup_args(q)
    d = mpc_batch(1, N=20)
    yield "mpc", (d["P"], d["c"][0], d["A"], d["b"][0], None, None, None, d["x_l"][0], d["x_u"][0])
    q = sparse_strongly_convex_qp(60, 20, 30, 0.1, seed=3)
    yield "random", (q["P"], q["c"], q["A"], q["b"], q["G"], q["h_l"], q["h_u"], q["x_l"], q["x_u"])


@pytest.mark.parametrize("name,args", list(_cases()))
def test_symbolic_matches_oracle_under_same_permutation(oracle, name, args):
    o = oracle.SparseSolver(oracle.default_settings(kkt_solver="sparse_ldlt")); o.setup(*args)
    P, AT, GT = o.scaled_matrices()
    mine = _symbolic(P, AT, GT)
    nk = P.shape[0] + AT.shape[1] + GT.shape[1]
    assert sorted(mine["perm"].tolist()) == list(range(nk))
    # nnz(KKT upper) = nnz(P_utri incl. every diagonal) + nnz(A) + nnz(G) + p + m   (kkt_full.hpp:39-170)
    Pu = sp.triu(sp.csc_matrix(P)); diag_present = int((Pu.tocoo().row == Pu.tocoo().col).sum())
    assert mine["nnz_kkt"] == Pu.nnz + (P.shape[0] - diag_present) + AT.nnz + GT.nnz + AT.shape[1] + GT.shape[1]
    # same permutation -> the oracle's LDLt::factorize_symbolic must count the same nnz(L) and flops
    o2 = oracle.SparseSolver(oracle.default_settings(kkt_solver="sparse_ldlt"), kkt_perm=mine["perm"]); o2.setup(*args)
    nnzL, flops = o2.ldlt_stats()
    assert nnzL == mine["nnz_L"]
    assert flops == pytest.approx(mine["flops"], rel=1e-12)
    # ordering quality: own minimum degree is in the same class as the oracle's
    nnzL_oracle, _ = o.ldlt_stats()
    assert mine["nnz_L"] <= 1.3 * nnzL_oracle + 16
    # user permutation: the backend composes it with a postorder of the elimination tree (supernodes become runs of
    # consecutive columns); that is an equivalent ordering -- same nnz(L) and flops as the user's -- and is idempotent
    rev = np.arange(nk)[::-1].copy()
    again = _symbolic(P, AT, GT, perm=rev)
    assert sorted(again["perm"].tolist()) == list(range(nk))
    o3 = oracle.SparseSolver(oracle.default_settings(kkt_solver="sparse_ldlt"), kkt_perm=rev); o3.setup(*args)
    assert o3.ldlt_stats()[0] == again["nnz_L"]
    assert np.array_equal(_symbolic(P, AT, GT, perm=again["perm"])["perm"], again["perm"])
    assert 1 <= again["levels"] <= nk


def test_symbolic_rejects_bad_permutation(oracle):
    q = sparse_strongly_convex_qp(10, 3, 4, 0.3, seed=1)
    P, AT, GT = sp.triu(q["P"]), sp.csc_matrix(q["A"]).T, sp.csc_matrix(q["G"]).T
    bad = np.zeros(17, dtype=np.int32)
    with pytest.raises(RuntimeError, match="invalid permutation"):
        _symbolic(P, AT, GT, perm=bad)


COND = {"sparse_ldlt_eq_cond": 1, "sparse_ldlt_ineq_cond": 2, "sparse_ldlt_cond": 3}


@pytest.mark.parametrize("solver", list(COND))
@pytest.mark.parametrize("name,args", list(_cases()))
def test_condensed_mode_symbolic_matches_oracle(oracle, name, args, solver):
    """KKT pattern of the condensed modes (kkt_{eq,ineq,all}_eliminated.hpp create_kkt_matrix: structural union of P, I,
    A^T A, G^T G in the top-left block) -> same nnz(KKT), and same nnz(L) / flops as the oracle under the same permutation"""
    from piqp_b200.backend import sparse_ldlt_symbolic
    mode = COND[solver]
    o = oracle.SparseSolver(oracle.default_settings(kkt_solver=solver)); o.setup(*args)
    P, AT, GT = o.scaled_matrices()
    n, p, m = P.shape[0], AT.shape[1], GT.shape[1]
    A = sp.csc_matrix(AT.T) if p else None
    G = sp.csc_matrix(GT.T) if m else None
    mine = sparse_ldlt_symbolic(P, A, G, mode=mode)
    nk = n + (0 if mode & 1 else p) + (0 if mode & 2 else m)
    assert sorted(mine["perm"].tolist()) == list(range(nk))
    # structural pattern of the top-left block
    pat = lambda M: sp.csc_matrix((np.ones(M.nnz), M.indices, M.indptr), shape=M.shape)
    Pu = pat(sp.csc_matrix(sp.triu(sp.csc_matrix(P))))
    top = Pu + sp.identity(n, format="csc")
    if mode & 1 and p:
        top = top + sp.triu(pat(sp.csc_matrix(AT)) @ pat(sp.csc_matrix(AT)).T)
    if mode & 2 and m:
        top = top + sp.triu(pat(sp.csc_matrix(GT)) @ pat(sp.csc_matrix(GT)).T)
    expect = sp.csc_matrix(top).nnz + (0 if mode & 1 else AT.nnz + p) + (0 if mode & 2 else GT.nnz + m)
    assert mine["nnz_kkt"] == expect
    o2 = oracle.SparseSolver(oracle.default_settings(kkt_solver=solver), kkt_perm=mine["perm"]); o2.setup(*args)
    nnzL, flops = o2.ldlt_stats()
    assert nnzL == mine["nnz_L"] and flops == pytest.approx(mine["flops"], rel=1e-12)
    assert mine["nnz_L"] <= 1.3 * o.ldlt_stats()[0] + 16
    assert mine["supernodes"] >= 1 and 1 <= mine["largest_front"] <= nk


def test_approximate_degree_ordering_is_as_good_as_exact_and_scales(monkeypatch):
    """the quotient-graph approximate minimum degree (the product's default) against the exact-external-degree version
    kept for cross-checks: fill within a few percent, and a 6 000-variable KKT is ordered in seconds"""
    import time
    from piqp_b200.backend import sparse_ldlt_symbolic
    q = sparse_strongly_convex_qp(600, 200, 300, 0.01, seed=9)
    amd = sparse_ldlt_symbolic(q["P"], q["A"], q["G"])
    monkeypatch.setenv("B200_ORDERING", "exact")
    exact = sparse_ldlt_symbolic(q["P"], q["A"], q["G"])
    monkeypatch.delenv("B200_ORDERING")
    assert amd["nnz_L"] <= 1.1 * exact["nnz_L"]
    q = sparse_strongly_convex_qp(3000, 1500, 1500, 0.005, seed=9)
    t0 = time.time()
    big = sparse_ldlt_symbolic(q["P"], q["A"], q["G"])
    assert time.time() - t0 < 30.0
    assert sorted(big["perm"].tolist()) == list(range(6000))


@pytest.mark.parametrize("mode", [0, 1, 2, 3])
def test_fast_column_counts_equal_the_reference_walk(monkeypatch, mode):
    """The O(nnz(K) alpha) elimination tree + skeleton column counts (Liu / Gilbert-Ng-Peyton) against the reference's O(nnz(L))
    row-subtree walk (sparse/ldlt.hpp:42-99, kept under B200_SYMBOLIC_WALK=1): identical permutation, nnz(L), flops, supernodes,
    largest front and level count on random KKT patterns of every KKTMode (the amalgamation pass runs on top of both)."""
    import piqp_b200
    rng = np.random.default_rng(5)
    for trial in range(6):
        n = int(rng.integers(30, 400)); p = int(rng.integers(0, n // 2)); m = int(rng.integers(0, n))
        q = sparse_strongly_convex_qp(n, p, m, float(rng.uniform(0.01, 0.15)), seed=100 + trial, eig_shift="gershgorin")
        P, A, G = sp.csc_matrix(q["P"]), (sp.csc_matrix(q["A"]) if p else None), (sp.csc_matrix(q["G"]) if m else None)
        monkeypatch.delenv("B200_SYMBOLIC_WALK", raising=False)
        fast = piqp_b200.sparse_ldlt_symbolic(P, A, G, mode=mode)
        monkeypatch.setenv("B200_SYMBOLIC_WALK", "1")
        walk = piqp_b200.sparse_ldlt_symbolic(P, A, G, mode=mode)
        for k in ("nnz_kkt", "nnz_L", "levels", "flops", "supernodes", "largest_front"):
            assert fast[k] == walk[k], (trial, mode, k, fast[k], walk[k])
        assert np.array_equal(fast["perm"], walk["perm"])
