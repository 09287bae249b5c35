"""GPU parity tests of the condensed sparse KKT modes (sparse_ldlt_eq_cond / _ineq_cond / _cond = KKTMode 1 / 2 / 3,
include/piqp/sparse/kkt_{eq,ineq,all}_eliminated.hpp) against the CPU oracle's restatement, all through the C-ABI.
Mirrors the reference's TEST_P over the sparse backends (tests/src/sparse/solver_test.cpp:443-451) and its
kkt_{eq,ineq,all}_eliminated_test.cpp factor/solve checks."""
import numpy as np
import pytest
import scipy.sparse as sp

from helpers import (dual_infeasible_qp, load_scenario_mpc, primal_infeasible_qp, setup_args, simple_qp, simple_qp_update)
from piqp_b200.synth import sparse_strongly_convex_qp

pytestmark = pytest.mark.gpu

MODES = {"sparse_ldlt_eq_cond": 1, "sparse_ldlt_ineq_cond": 2, "sparse_ldlt_cond": 3}


def _rel(a, b):
    return np.abs(a - b).max() / max(1.0, np.abs(b).max())


def _vtable(oracle, b200):
    from piqp_b200.backend import c_abi_vtable
    vt = oracle.BackendVTable()
    for k, v in c_abi_vtable().items():
        setattr(vt, k, v)
    return vt


def _random_args(n=80, p=25, m=40, seed=3, sparsity=0.08):
    q = sparse_strongly_convex_qp(n, p, m, sparsity, seed=seed)
    return (q["P"], q["c"], q["A"], q["b"], q["G"], q["h_l"], q["h_u"], q["x_l"], q["x_u"])


def _dense_kkt(P, AT, GT, delta, x_reg, z_reg):
    """the FULL 3x3 system every mode must solve (kkt_solver_base.hpp:34)"""
    n, p, m = P.shape[0], AT.shape[1], GT.shape[1]
    Pf = P + sp.triu(P, 1).T
    return sp.bmat([[Pf + sp.diags(x_reg), AT, GT], [AT.T, -delta * sp.eye(p), None], [GT.T, None, -sp.diags(z_reg)]]).tocsc(), n, p, m


@pytest.mark.parametrize("kernels", ["frontal", "frontal_hbm_fronts", "levels"])
@pytest.mark.parametrize("case", ["notebook", "random", "no_eq", "no_ineq"])
@pytest.mark.parametrize("solver", list(MODES))
def test_cond_backend_factor_solve_parity(oracle, b200, solver, case, kernels, monkeypatch):
    monkeypatch.setenv("B200_LDLT_LEVELS", "1" if kernels == "levels" else "0")
    monkeypatch.setenv("B200_FRONT_SMEM_ROWS", "6" if kernels == "frontal_hbm_fronts" else "0")
    if case == "notebook":
        q, _ = load_scenario_mpc(); args = setup_args(q)
    elif case == "random":
        args = _random_args()
    elif case == "no_eq":
        a = _random_args(40, 5, 30, seed=7); args = (a[0], a[1], None, None) + a[4:]
    else:
        a = _random_args(40, 15, 5, seed=8); args = a[:4] + (None, None, None, None, None)
    mode = MODES[solver]
    o = oracle.SparseSolver(oracle.default_settings(kkt_solver=solver)); o.setup(*args)
    P, AT, GT = o.scaled_matrices()
    n, p, m = o.dims[:3]
    be = b200.SparseKKT(P, AT, GT, mode=mode)
    info = be.symbolic_info()
    nk = n + (0 if mode & 1 else p) + (0 if mode & 2 else m)
    assert be.n_kkt == nk and sorted(info["perm"].tolist()) == list(range(nk))
    o2 = oracle.SparseSolver(oracle.default_settings(kkt_solver=solver), kkt_perm=info["perm"]); o2.setup(*args)
    assert o2.ldlt_stats()[0] == info["nnz_L"]
    rng = np.random.default_rng(0)
    for trial in range(2):
        x_reg = rng.uniform(0.5, 1.5, n); z_reg = rng.uniform(0.5, 2.0, m); delta = float(rng.uniform(0.5, 1.5))
        assert o.backend_factor(delta, x_reg, z_reg) == 1 and o2.backend_factor(delta, x_reg, z_reg) == 1
        assert be.update_scalings_and_factor(delta, x_reg, z_reg) is True
        rx, ry, rz = rng.standard_normal(n), rng.standard_normal(p), rng.standard_normal(m)
        lg = be.solve(rx, ry, rz)
        for ref, tol in ((o2.backend_solve(rx, ry, rz), 1e-11), (o.backend_solve(rx, ry, rz), 1e-9)):
            for a, b in zip(lg, ref):
                if len(b):
                    assert _rel(a, b) < tol
        # size-independent property: the condensed solve satisfies the FULL 3x3 system
        K, *_ = _dense_kkt(sp.csc_matrix(P), sp.csc_matrix(AT), sp.csc_matrix(GT), delta, x_reg, z_reg)
        sol = np.concatenate(lg)
        assert np.abs(K @ sol - np.concatenate([rx, ry, rz])).max() < 1e-9 * max(1.0, np.abs(sol).max())
    cl = be.clone()
    for a, b in zip(cl.solve(rx, ry, rz), be.solve(rx, ry, rz)):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("solver", list(MODES))
def test_cond_update_data_equals_fresh(oracle, b200, solver):
    """update_data_impl of the condensed modes (A^T A is recomputed on KKT_UPDATE_A): incremental == fresh, exactly"""
    args = _random_args(50, 10, 20, seed=11)
    o = oracle.SparseSolver(oracle.default_settings(kkt_solver=solver)); o.setup(*args)
    P, AT, GT = o.scaled_matrices()
    be = b200.SparseKKT(P, AT, GT, mode=MODES[solver])
    n, p, m = o.dims[:3]
    rng = np.random.default_rng(2)
    P2, AT2, GT2 = sp.csc_matrix(sp.triu(P)) * 1.2, AT.copy(), GT.copy()
    AT2.data = AT2.data * rng.uniform(0.8, 1.3, AT2.nnz); GT2.data = GT2.data * 0.7
    be.update_data(7, P2, AT2, GT2)
    fresh = b200.SparseKKT(P2, AT2, GT2, perm=be.symbolic_info()["perm"], mode=MODES[solver])
    x_reg = rng.uniform(0.5, 1.5, n); z_reg = rng.uniform(0.5, 2.0, m)
    assert be.update_scalings_and_factor(0.8, x_reg, z_reg) and fresh.update_scalings_and_factor(0.8, x_reg, z_reg)
    r = (rng.standard_normal(n), rng.standard_normal(p), rng.standard_normal(m))
    for a, b in zip(be.solve(*r), fresh.solve(*r)):
        assert np.array_equal(a, b)
    K, *_ = _dense_kkt(P2, AT2, GT2, 0.8, x_reg, z_reg)
    sol = np.concatenate(be.solve(*r))
    assert np.abs(K @ sol - np.concatenate(r)).max() < 1e-9 * max(1.0, np.abs(sol).max())


@pytest.mark.parametrize("case", ["notebook", "random"])
@pytest.mark.parametrize("solver", list(MODES))
def test_reference_style_solver_drives_cuda_cond_backend(oracle, b200, solver, case):
    """oracle KKTSystem + IP loop -> b200kkt_sparse_create(mode) through the C-ABI table: same iterations and solution"""
    if case == "notebook":
        q, g = load_scenario_mpc(); args = setup_args(q)
    else:
        args = _random_args(60, 20, 30, seed=5)
    st = oracle.default_settings(kkt_solver=solver)
    cpu = oracle.SparseSolver(st); cpu.setup(*args); assert cpu.solve() == 1
    gpu = oracle.SparseSolver(st, backend_vtable=_vtable(oracle, b200)); gpu.setup(*args); assert gpu.solve() == 1
    rc, rg = cpu.result(), gpu.result()
    assert rg.info.iter == rc.info.iter
    assert np.abs(rg.x - rc.x).max() <= 1e-8 * max(1.0, np.abs(rc.x).max())


@pytest.mark.parametrize("solver", list(MODES))
def test_batched_cond_matches_oracle_known_answers_and_statuses(oracle, b200, solver):
    """device-resident IP loop over the condensed backends: random QPs == oracle (status, iterations, x), the reference's
    known-answer QP + its update (sparse/solver_test.cpp:67-107) and the infeasibility statuses"""
    B = 4
    base = sparse_strongly_convex_qp(70, 20, 35, 0.08, seed=21)
    rng = np.random.default_rng(4)
    Pu = sp.csc_matrix(sp.triu(base["P"])); A = sp.csc_matrix(base["A"]); G = sp.csc_matrix(base["G"])
    Pu.sort_indices(); A.sort_indices(); G.sort_indices()
    Ax = np.stack([A.data * rng.uniform(0.8, 1.2, A.nnz) for _ in range(B)])
    Gx = np.stack([G.data * rng.uniform(0.8, 1.2, G.nnz) for _ in range(B)])
    c = np.stack([base["c"] + 0.1 * rng.standard_normal(70) for _ in range(B)])
    s = b200.SparseSolverBatched(kkt_solver=solver)
    st = lambda v: np.broadcast_to(v, (B, len(v)))
    s.setup(B, Pu, c, A, st(base["b"]), G, st(base["h_l"]), st(base["h_u"]), st(base["x_l"]), st(base["x_u"]), Ax=Ax, Gx=Gx)
    infos = s.solve(); r = s.result()
    for k in range(B):
        Ak = sp.csc_matrix((Ax[k], A.indices, A.indptr), shape=A.shape)
        Gk = sp.csc_matrix((Gx[k], G.indices, G.indptr), shape=G.shape)
        o = oracle.SparseSolver(oracle.default_settings(kkt_solver=solver))
        o.setup(Pu, c[k], Ak, base["b"], Gk, base["h_l"], base["h_u"], base["x_l"], base["x_u"])
        status = o.solve(); ro = o.result()
        assert infos[k].status == status == 1
        assert infos[k].iter == ro.info.iter, (k, infos[k].iter, ro.info.iter)
        assert np.abs(r.x[k] - ro.x).max() <= 1e-8 * max(1.0, np.abs(ro.x).max())
    q1 = simple_qp(); q2 = simple_qp_update(q1)
    S = lambda M: sp.csc_matrix(M)
    t = b200.SparseSolverBatched(kkt_solver=solver)
    P1, A1, G1, P2, A2 = S(q1["P"]), S(q1["A"]), S(q1["G"]), S(q2["P"]), S(q2["A"])
    t.setup(1, P1, q1["c"], A1, q1["b"], G1, q1["h_l"], q1["h_u"], q1["x_l"], q1["x_u"])
    assert t.solve()[0].status == 1
    r = t.result()
    assert np.allclose(r.x[0], [0.4285714, 0.2142857], atol=1e-6) and abs(r.y[0, 0] + 1.5714286) < 1e-6
    t.update(Px=P2.data[None], c=q2["c"][None], Ax=A2.data[None], b=q2["b"][None], h_u=q2["h_u"][None], x_u=q2["x_u"][None])
    assert t.solve()[0].status == 1
    assert np.allclose(t.result().x[0], [0.2763157, 0.0921056], atol=1e-6)
    for make, status in ((primal_infeasible_qp, -2), (dual_infeasible_qp, -3)):
        q = make()
        u = b200.SparseSolverBatched(kkt_solver=solver)
        u.setup(1, S(q["P"]), q["c"], S(q["A"]) if q.get("A") is not None else None, q.get("b"), S(q["G"]) if q.get("G") is not None else None,
                q.get("h_l"), q.get("h_u"), q.get("x_l"), q.get("x_u"))
        assert u.solve()[0].status == status
