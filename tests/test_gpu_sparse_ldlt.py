"""GPU parity tests of the general sparse LDL^T CUDA backend (sparse_ldlt, KKTMode FULL) against the CPU oracle's
restatement of sparse::KKT / LDLt and against the trace the real reference printed, all through the C-ABI."""
import numpy as np
import pytest
import scipy.sparse as sp

from helpers import (dual_infeasible_qp, load_scenario_mpc, primal_infeasible_qp, setup_args, simple_qp, simple_qp_update,
                     trace_as_printed)
from piqp_b200.synth import mpc_batch, sparse_strongly_convex_qp

pytestmark = pytest.mark.gpu


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


def _select_kernels(monkeypatch, kernels):
    """kernel families of the sparse_ldlt backend: CTA-per-QP multifrontal (default), the same with fronts forced into HBM,
    level-scheduled simplicial, and the whole-GPU ("wide") schedule with default / tiny thresholds"""
    monkeypatch.setenv("B200_LDLT_LEVELS", "1" if kernels == "levels" else "0")
    monkeypatch.setenv("B200_FRONT_SMEM_ROWS", "6" if kernels in ("frontal_hbm_fronts", "wide_tiny") else "0")   # fronts > 6 rows -> HBM-front path
    monkeypatch.setenv("B200_LDLT_WIDE", "1" if kernels.startswith("wide") else "0")
    if kernels == "wide_tiny":
        monkeypatch.setenv("B200_WIDE_WS", "2"); monkeypatch.setenv("B200_WIDE_SB", "8")
    else:
        monkeypatch.delenv("B200_WIDE_WS", raising=False); monkeypatch.delenv("B200_WIDE_SB", raising=False)


@pytest.mark.parametrize("kernels", ["frontal", "frontal_hbm_fronts", "levels", "wide", "wide_tiny"])
@pytest.mark.parametrize("case", ["notebook", "mpc", "random", "no_eq", "no_ineq"])
@pytest.mark.parametrize("own_perm", [True, False])
def test_backend_factor_solve_eval_parity(oracle, b200, case, own_perm, kernels, monkeypatch):
    """sparse/kkt_test style (tests/src/sparse/kkt_*_test.cpp): same rho/delta/scalings -> same solve / mat-vec results.
    All numeric kernel families (see _select_kernels)."""
    _select_kernels(monkeypatch, kernels)
    if case == "notebook":
        q, _ = load_scenario_mpc(); args = setup_args(q)
    elif case == "mpc":
        d = mpc_batch(1, N=15); args = (d["P"], d["c"][0], d["A"], d["b"][0], None, None, None, d["x_l"][0], d["x_u"][0])
    elif case == "random":
        args = _random_args()
    elif case == "no_eq":
        a = _random_args(40, 5, 30, seed=7); args = (a[0], a[1], None, None) + a[4:]
    else:
        a = _random_args(40, 15, 5, seed=8); args = a[:4] + (None, None, None, None, None)
    o = oracle.SparseSolver(oracle.default_settings(kkt_solver="sparse_ldlt")); o.setup(*args)
    P, AT, GT = o.scaled_matrices()
    n, p, m = o.dims[:3]
    nk = n + p + m
    perm = None if own_perm else np.random.default_rng(1).permutation(nk).astype(np.int32)
    be = b200.SparseKKT(P, AT, GT, perm=perm)
    info = be.symbolic_info()
    assert sorted(info["perm"].tolist()) == list(range(nk))
    # (a user permutation is composed with a postorder of the elimination tree: an equivalent ordering)
    # oracle with the SAME permutation: identical elimination order -> agreement to rounding
    o2 = oracle.SparseSolver(oracle.default_settings(kkt_solver="sparse_ldlt"), kkt_perm=info["perm"]); o2.setup(*args)
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
        x, y, z = rng.standard_normal(n), rng.standard_normal(p), rng.standard_normal(m)
        assert _rel(be.eval_P_x(0.7, x), o.backend_eval_P_x(0.7, x)) < 1e-12
        for a, b in zip(be.eval_A_xn_and_AT_xt(-1.0, 1.0, x, y), o.backend_eval_A(-1.0, 1.0, x, y)):
            if len(b):
                assert _rel(a, b) < 1e-12
        for a, b in zip(be.eval_G_xn_and_GT_xt(1.0, 1.0, x, z), o.backend_eval_G(1.0, 1.0, x, z)):
            if len(b):
                assert _rel(a, b) < 1e-12
    cl = be.clone()
    for a, b in zip(cl.solve(rx, ry, rz), be.solve(rx, ry, rz)):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("mode", [0, 3])
@pytest.mark.parametrize("sb", [128, 32])
def test_wide_schedule_mid_size_multi_panel(oracle, b200, mode, sb, monkeypatch):
    """whole-GPU schedule on a problem whose root front spans several 64-column panels, 128 x 128 DMMA tiles and solve blocks
    (BASELINE config 3 family at n_kkt = 700): agreement with the oracle under the same permutation, residual of the full
    3x3 system, determinism of a clone; update matrices in per-supernode slots, pull-form extend-add"""
    monkeypatch.setenv("B200_LDLT_WIDE", "1"); monkeypatch.setenv("B200_LDLT_LEVELS", "0"); monkeypatch.setenv("B200_FRONT_SMEM_ROWS", "0")
    monkeypatch.setenv("B200_WIDE_SB", str(sb)); monkeypatch.delenv("B200_WIDE_WS", raising=False)
    q = sparse_strongly_convex_qp(350, 150, 200, 0.03, seed=17)
    solver = {0: "sparse_ldlt", 3: "sparse_ldlt_cond"}[mode]
    o = oracle.SparseSolver(oracle.default_settings(kkt_solver=solver))
    o.setup(q["P"], q["c"], q["A"], q["b"], q["G"], q["h_l"], q["h_u"], q["x_l"], q["x_u"])
    P, AT, GT = o.scaled_matrices()
    n, p, m = o.dims[:3]
    be = b200.SparseKKT(P, AT, GT, mode=mode)
    info = be.symbolic_info()
    o2 = oracle.SparseSolver(oracle.default_settings(kkt_solver=solver), kkt_perm=info["perm"])
    o2.setup(q["P"], q["c"], q["A"], q["b"], q["G"], q["h_l"], q["h_u"], q["x_l"], q["x_u"])
    rng = np.random.default_rng(0)
    x_reg = rng.uniform(0.5, 1.5, n); z_reg = rng.uniform(0.5, 2.0, m); delta = 0.9
    assert o2.backend_factor(delta, x_reg, z_reg) == 1 and be.update_scalings_and_factor(delta, x_reg, z_reg) is True
    r = (rng.standard_normal(n), rng.standard_normal(p), rng.standard_normal(m))
    lg = be.solve(*r)
    for a, b in zip(lg, o2.backend_solve(*r)):
        assert _rel(a, b) < 1e-10
    Pf = sp.csc_matrix(P) + sp.triu(sp.csc_matrix(P), 1).T
    K = sp.bmat([[Pf + sp.diags(x_reg), AT, GT], [AT.T, -delta * sp.eye(p), None], [GT.T, None, -sp.diags(z_reg)]]).tocsc()
    sol = np.concatenate(lg)
    assert np.abs(K @ sol - np.concatenate(r)).max() < 1e-9 * max(1.0, np.abs(sol).max())
    cl = be.clone()
    for a, b in zip(cl.solve(*r), be.solve(*r)):
        assert np.array_equal(a, b)
    assert be.update_scalings_and_factor(delta, x_reg, z_reg) is True       # refactor: bitwise the same factor
    for a, b in zip(be.solve(*r), lg):
        assert np.array_equal(a, b)


def test_update_data_refreshes_values(oracle, b200):
    """KKT::update_data (sparse/kkt.hpp:72-81, kkt_full.hpp:212-251): new P/A/G values, same pattern"""
    args = _random_args(50, 10, 20, seed=11)
    o = oracle.SparseSolver(oracle.default_settings(kkt_solver="sparse_ldlt")); o.setup(*args)
    P, AT, GT = o.scaled_matrices()
    be = b200.SparseKKT(P, AT, GT)
    n, p, m = o.dims[:3]
    rng = np.random.default_rng(2)
    P2, AT2, GT2 = P.copy(), AT.copy(), GT.copy()
    P2.data = P2.data * rng.uniform(0.9, 1.1, P2.nnz); AT2.data = AT2.data * 1.3; GT2.data = GT2.data * 0.7
    P2 = P2 + sp.diags(np.full(n, 0.5))     # keeps the pattern when the diagonal is present ...
    P2 = sp.csc_matrix(sp.triu(P2))
    if P2.nnz != sp.triu(P).nnz:              # ... otherwise fall back to a pure rescale
        P2 = sp.csc_matrix(sp.triu(P)) * 1.2
    be.update_data(7, P2, AT2, GT2)
    fresh = b200.SparseKKT(P2, AT2, GT2, perm=be.symbolic_info()["perm"])
    x_reg = rng.uniform(0.5, 1.5, n); z_reg = rng.uniform(0.5, 2.0, m)
    assert be.update_scalings_and_factor(0.8, x_reg, z_reg) and fresh.update_scalings_and_factor(0.8, x_reg, z_reg)
    r = (rng.standard_normal(n), rng.standard_normal(p), rng.standard_normal(m))
    for a, b in zip(be.solve(*r), fresh.solve(*r)):
        assert np.array_equal(a, b)
    K = sp.bmat([[P2 + sp.triu(P2, 1).T + sp.diags(x_reg), AT2, GT2], [AT2.T, -0.8 * sp.eye(p), None], [GT2.T, None, -sp.diags(z_reg)]]).tocsc()
    sol = np.concatenate(be.solve(*r))
    assert np.abs(K @ sol - np.concatenate(r)).max() < 1e-9 * max(1.0, np.abs(sol).max())


@pytest.mark.parametrize("piece", ["Px", "Ax", "Gx"])
@pytest.mark.parametrize("solver", ["sparse_ldlt", "sparse_ldlt_eq_cond", "sparse_ldlt_cond"])
def test_batched_partial_matrix_update_equals_fresh_setup(b200, solver, piece):
    """update() of ONE matrix with preconditioner_reuse_on_update = 0: Ruiz is recomputed, every matrix is rescaled, so every
    cached copy (permuted KKT values, A^T A of the condensed modes) must be refreshed; bar: a fresh setup()"""
    from piqp_b200.synth import sparse_batch
    B = 3
    d = sparse_batch(B, 60, 15, 30, 0.08, seed0=7)
    def run(vals, s=None):
        if s is None:
            s = b200.SparseSolverBatched(kkt_solver=solver)
            s.setup(B, d["P"], d["c"], d["A"], d["b"], d["G"], d["h_l"], d["h_u"], d["x_l"], d["x_u"], Px=vals["Px"], Ax=vals["Ax"], Gx=vals["Gx"])
        infos = s.solve()
        return s, infos, s.result()
    v0 = {k: d[k] for k in ("Px", "Ax", "Gx")}
    s, _, _ = run(v0)
    rng = np.random.default_rng(3)
    v1 = dict(v0)
    if piece == "Px":
        diag = np.repeat(np.arange(60), np.diff(d["P"].indptr)) == d["P"].indices
        v1["Px"] = v0["Px"] * np.where(diag, rng.uniform(1.5, 3.0, diag.shape), 1.0)
    else:
        v1[piece] = v0[piece] * rng.uniform(0.5, 3.0, v0[piece].shape)
    s.update(**{piece: v1[piece]})
    _, iu, ru = run(v1, s)
    _, i_f, rf = run(v1)
    for b in range(B):
        assert iu[b].status == i_f[b].status
        assert iu[b].iter == i_f[b].iter, (b, iu[b].iter, i_f[b].iter)
        if i_f[b].status == 1:
            assert _rel(ru.x[b], rf.x[b]) <= 1e-8


@pytest.mark.parametrize("case", ["notebook", "random"])
def test_reference_style_solver_drives_cuda_sparse_ldlt(oracle, b200, case):
    """oracle KKTSystem + IP loop -> b200kkt_sparse_* through the C-ABI table: same iterations and solution"""
    if case == "notebook":
        q, g = load_scenario_mpc(); args = setup_args(q)
    else:
        args = _random_args(60, 20, 30, seed=5)
    st = oracle.default_settings(kkt_solver="sparse_ldlt")
    cpu = oracle.SparseSolver(st); cpu.setup(*args); assert cpu.solve() == 1
    gpu = oracle.SparseSolver(st, backend_vtable=_vtable(oracle, b200)); gpu.setup(*args); assert gpu.solve() == 1
    rc, rg = cpu.result(), gpu.result()
    assert rg.info.iter == rc.info.iter
    if case == "notebook":
        assert rg.info.iter == g["iterations"]
    assert np.abs(rg.x - rc.x).max() <= 1e-8 * max(1.0, np.abs(rc.x).max())


def test_batched_notebook_golden_trace_sparse_ldlt(oracle, b200):
    """device-resident IP loop + sparse LDL^T CUDA backend reproduce the trace the REAL reference printed for sparse_ldlt"""
    q, g = load_scenario_mpc()
    s = b200.SparseSolverBatched(kkt_solver="sparse_ldlt")
    s.settings.verbose = 2
    s.setup(2, q["P"], q["c"], q["A"], q["b"], None, None, None, q["x_l"], q["x_u"])
    infos = s.solve()
    assert [i.status for i in infos] == [1, 1] and [i.iter for i in infos] == [g["iterations"]] * 2
    mine = trace_as_printed(s.trace(0)); golden = np.array(g["trace_sparse_ldlt"])[:, 1:]
    assert mine.shape == golden.shape
    k = 12
    assert np.all(np.abs(mine[:k, :5] - golden[:k, :5]) <= 3e-5 * np.abs(golden[:k, :5]))
    assert np.abs(mine[:, 8:] - golden[:, 8:]).max() <= 6e-5
    r = s.result()
    assert np.array_equal(r.x[0], r.x[1])


def test_batched_random_sparse_matches_oracle(oracle, b200):
    """BASELINE config 3 family (random sparse strongly convex QPs) at small size: per-instance values, shared pattern"""
    B = 5
    base = sparse_strongly_convex_qp(70, 20, 35, 0.08, seed=21)
    rng = np.random.default_rng(4)
    Pu = sp.csc_matrix(sp.triu(base["P"])); A = sp.csc_matrix(base["A"]); G = sp.csc_matrix(base["G"])
    Pu.sort_indices(); A.sort_indices(); G.sort_indices()
    diag_mask = (Pu.tocoo().row == Pu.tocoo().col)
    Px = np.stack([Pu.data * np.where(diag_mask, rng.uniform(1.0, 1.2), 1.0) for _ in range(B)])
    Ax = np.stack([A.data * rng.uniform(0.8, 1.2, A.nnz) for _ in range(B)])
    Gx = np.stack([G.data * rng.uniform(0.8, 1.2, G.nnz) for _ in range(B)])
    c = np.stack([base["c"] + 0.1 * rng.standard_normal(70) for _ in range(B)])
    s = b200.SparseSolverBatched(kkt_solver="sparse_ldlt")
    st = lambda v: np.broadcast_to(v, (B, len(v)))
    s.setup(B, Pu, c, A, st(base["b"]), G, st(base["h_l"]), st(base["h_u"]), st(base["x_l"]), st(base["x_u"]), Px=Px, Ax=Ax, Gx=Gx)
    infos = s.solve(); r = s.result()
    for k in range(B):
        Pk = sp.csc_matrix((Px[k], Pu.indices, Pu.indptr), shape=Pu.shape)
        Ak = sp.csc_matrix((Ax[k], A.indices, A.indptr), shape=A.shape)
        Gk = sp.csc_matrix((Gx[k], G.indices, G.indptr), shape=G.shape)
        o = oracle.SparseSolver(oracle.default_settings(kkt_solver="sparse_ldlt"))
        o.setup(Pk, c[k], Ak, base["b"], Gk, base["h_l"], base["h_u"], base["x_l"], base["x_u"])
        status = o.solve(); ro = o.result()
        assert infos[k].status == status == 1
        assert infos[k].iter == ro.info.iter, (k, infos[k].iter, ro.info.iter)
        assert np.abs(r.x[k] - ro.x).max() <= 1e-8 * max(1.0, np.abs(ro.x).max())


def test_batched_sparse_ldlt_known_answers_infeasibility_and_update(oracle, b200):
    """sparse/solver_test.cpp:67-107 golden values + the infeasibility statuses through kkt_solver = sparse_ldlt"""
    q1 = simple_qp(); q2 = simple_qp_update(q1)
    S = lambda M: sp.csc_matrix(M)
    s = b200.SparseSolverBatched(kkt_solver="sparse_ldlt")
    P1, A1, G1 = S(q1["P"]), S(q1["A"]), S(q1["G"])
    P2, A2 = S(q2["P"]), S(q2["A"])
    st = lambda a, b: np.stack([a, b])
    s.setup(2, P1, st(q1["c"], q2["c"]), A1, st(q1["b"], q2["b"]), G1, st(q1["h_l"], q2["h_l"]), st(q1["h_u"], q2["h_u"]),
            st(q1["x_l"], q2["x_l"]), st(q1["x_u"], q2["x_u"]), Px=st(P1.data, P2.data), Ax=st(A1.data, A2.data))
    infos = s.solve(); r = s.result()
    assert [i.status for i in infos] == [1, 1]
    assert np.allclose(r.x[0], [0.4285714, 0.2142857], atol=1e-6) and abs(r.y[0, 0] + 1.5714286) < 1e-6
    assert np.allclose(r.x[1], [0.2763157, 0.0921056], atol=1e-6) and abs(r.y[1, 0] + 1.2105263) < 1e-6
    s.update(Px=st(P2.data, P2.data), c=st(q2["c"], q2["c"]), Ax=st(A2.data, A2.data), b=st(q2["b"], q2["b"]), h_u=st(q2["h_u"], q2["h_u"]), x_u=st(q2["x_u"], q2["x_u"]))
    infos = s.solve(); r = s.result()
    assert [i.status for i in infos] == [1, 1]
    assert np.allclose(r.x[0], [0.2763157, 0.0921056], atol=1e-6) and np.allclose(r.x[0], r.x[1], atol=1e-9)
    for make, status in ((primal_infeasible_qp, -2), (dual_infeasible_qp, -3)):
        q = make()
        t = b200.SparseSolverBatched(kkt_solver="sparse_ldlt")
        t.setup(1, S(q["P"]), q["c"], S(q["A"]) if q.get("A") is not None else None, q.get("b"), S(q["G"]) if q.get("G") is not None else None,
                q.get("h_l"), q.get("h_u"), q.get("x_l"), q.get("x_u"))
        assert t.solve()[0].status == status


@pytest.mark.parametrize("solver", ["sparse_ldlt", "sparse_ldlt_cond"])
def test_batched_ip_loop_over_the_wide_schedule(oracle, b200, solver, monkeypatch):
    """device-resident IP loop (batch of 3, per-instance values, instances retire at different iterations -> `active` masks)
    over the whole-GPU schedule with tiny thresholds: every wide kernel sees batch strides and inactive instances"""
    monkeypatch.setenv("B200_LDLT_WIDE", "1"); monkeypatch.setenv("B200_LDLT_LEVELS", "0"); monkeypatch.setenv("B200_FRONT_SMEM_ROWS", "6")
    monkeypatch.setenv("B200_WIDE_WS", "2"); monkeypatch.setenv("B200_WIDE_SB", "8")
    B = 3
    base = sparse_strongly_convex_qp(70, 20, 35, 0.08, seed=21)
    rng = np.random.default_rng(4)
    Pu = sp.csc_matrix(sp.triu(base["P"])); A = sp.csc_matrix(base["A"]); G = sp.csc_matrix(base["G"])
    Pu.sort_indices(); A.sort_indices(); G.sort_indices()
    Ax = np.stack([A.data * rng.uniform(0.8, 1.2, A.nnz) for _ in range(B)])
    Gx = np.stack([G.data * rng.uniform(0.8, 1.2, G.nnz) for _ in range(B)])
    c = np.stack([base["c"] + 0.3 * rng.standard_normal(70) for _ in range(B)])
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
