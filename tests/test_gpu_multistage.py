"""GPU parity tests of the multistage (block-tridiagonal-arrow) CUDA backend against the CPU oracle and against the
numbers the real reference prints in its documentation notebook, all through the C-ABI."""
import numpy as np
import pytest
import scipy.sparse as sp

from helpers import load_scenario_mpc, setup_args, simple_qp, simple_qp_update, trace_as_printed
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


def test_structure_detection_matches_reference_print_and_oracle(oracle, b200):
    """'block sizes: 8,6 8,6 8,6 14,0 (x3) / arrow width: 8' (notebook :550-551) from the product's own detection"""
    q, g = load_scenario_mpc()
    o = oracle.SparseSolver(oracle.default_settings(kkt_solver="sparse_multistage")); o.setup(*setup_args(q))
    P, AT, GT = o.scaled_matrices()
    be = b200.MultistageKKT(P, AT, GT)
    blocks = be.block_info()
    assert [[d, off] for (_, d, off) in blocks[:-1]] == g["multistage_block_sizes"]
    assert blocks[-1][1] == g["multistage_arrow_width"]
    assert blocks == o.multistage_blocks()
    d = mpc_batch(1, N=20)
    o = oracle.SparseSolver(oracle.default_settings(kkt_solver="sparse_multistage")); o.setup(d["P"], d["c"][0], d["A"], d["b"][0], None, None, None, d["x_l"][0], d["x_u"][0])
    P, AT, GT = o.scaled_matrices()
    assert b200.MultistageKKT(P, AT, GT).block_info() == o.multistage_blocks()


@pytest.mark.parametrize("N,segments", [(30, None), (30, "2"), (30, "7"), (30, "10"), (100, None), (100, "4"), (57, "12")])
def test_parallel_in_horizon_partition_matches_sequential_chain_and_oracle(oracle, b200, N, segments, monkeypatch):
    """the horizon cut into K runs at K-1 separator stages (runs factorised / substituted concurrently, spikes,
    reduced chain on the separators; multistage_partition.cuh) == the sequential warp chain (B200_MS_NO_PARTITION=1) == the oracle's
    factor_kkt / solve_llt_in_place restatement, for default and forced K, including runs of 1-2 stages"""
    d = mpc_batch(1, N=N)
    args = (d["P"], d["c"][0], d["A"], d["b"][0], None, None, None, d["x_l"][0], d["x_u"][0])
    o = oracle.SparseSolver(oracle.default_settings(kkt_solver="sparse_multistage")); o.setup(*args)
    P, AT, GT = o.scaled_matrices()
    n, p, m = o.dims[:3]
    rng = np.random.default_rng(N)
    x_reg = rng.uniform(0.5, 1.5, n); z_reg = np.zeros(m); delta = 0.7
    r = (rng.standard_normal(n), rng.standard_normal(p), np.zeros(m))
    assert o.backend_factor(delta, x_reg, z_reg) == 1
    ref = o.backend_solve(*r)
    out = {}
    for mode in ("partition", "sequential"):
        monkeypatch.setenv("B200_MS_NO_PARTITION", "1" if mode == "sequential" else "0")
        if segments and mode == "partition":
            monkeypatch.setenv("B200_MS_SEGMENTS", segments)
        else:
            monkeypatch.delenv("B200_MS_SEGMENTS", raising=False)
        be = b200.MultistageKKT(P, AT, GT)
        assert be.update_scalings_and_factor(delta, x_reg, z_reg) is True
        out[mode] = be.solve(*r)
        assert be.update_scalings_and_factor(0.3, x_reg * 2, z_reg) is True       # refactor with other scalings, then back: no stale state
        assert be.update_scalings_and_factor(delta, x_reg, z_reg) is True
        for a, b in zip(be.solve(*r), out[mode]):
            assert np.array_equal(a, b)
        cl = be.clone()
        for a, b in zip(cl.solve(*r), out[mode]):
            assert np.array_equal(a, b)
    for a, b, c in zip(out["partition"], out["sequential"], ref):
        if len(c):
            assert _rel(a, b) < 1e-10 and _rel(a, c) < 1e-9, (_rel(a, b), _rel(a, c))


@pytest.mark.parametrize("path", ["warp_chain", "generic"])
@pytest.mark.parametrize("case", ["notebook", "mpc", "random_with_G"])
def test_backend_factor_solve_eval_parity(oracle, b200, case, path, monkeypatch):
    """multistage_kkt_test.cpp:24-98 style: same rho/delta/scalings -> same solve and mat-vec results (vs the oracle's
    multistage AND vs its sparse_ldlt backend).  Both kernel families are exercised: the one-warp-per-QP chain kernels
    (fronts of <= 32 rows, multistage_chain.cuh) and the general shared-memory kernels (B200_MS_GENERIC=1)."""
    monkeypatch.setenv("B200_MS_GENERIC", "1" if path == "generic" else "0")
    if case == "notebook":
        q, _ = load_scenario_mpc(); args = setup_args(q)
    elif case == "mpc":
        d = mpc_batch(1, N=30); args = (d["P"], d["c"][0], d["A"], d["b"][0], None, None, None, d["x_l"][0], d["x_u"][0])
    else:
        # banded random QP with inequality rows so that G^T Z^-1 G contributes (block detection must cope with it)
        rng = np.random.default_rng(5); n = 60
        P = sp.diags([rng.uniform(1, 2, n), 0.3 * rng.standard_normal(n - 1)], [0, 1]).tocsc()
        A = sp.csc_matrix(sp.diags([rng.standard_normal(n - 2), rng.standard_normal(n - 2)], [0, 2], shape=(n - 2, n)))
        G = sp.csc_matrix(sp.diags([rng.standard_normal(n - 1), rng.standard_normal(n - 1)], [0, 1], shape=(n - 1, n)))
        args = (P, rng.standard_normal(n), A, rng.standard_normal(n - 2), G, -np.ones(n - 1) * 5, np.ones(n - 1) * 5, None, None)
    oms = oracle.SparseSolver(oracle.default_settings(kkt_solver="sparse_multistage")); oms.setup(*args)
    old = oracle.SparseSolver(oracle.default_settings(kkt_solver="sparse_ldlt")); old.setup(*args)
    P, AT, GT = oms.scaled_matrices()
    be = b200.MultistageKKT(P, AT, GT)
    n, p, m = oms.dims[:3]
    rng = np.random.default_rng(0)
    for trial in range(2):
        x_reg = rng.uniform(0.5, 1.5, n); z_reg = rng.uniform(0.5, 2.0, m); delta = float(rng.uniform(0.5, 1.5))
        assert oms.backend_factor(delta, x_reg, z_reg) == 1 and old.backend_factor(delta, x_reg, z_reg) == 1
        assert be.update_scalings_and_factor(delta, x_reg, z_reg) is True
        rx, ry, rz = rng.standard_normal(n), rng.standard_normal(p), rng.standard_normal(m)
        lg = be.solve(rx, ry, rz)
        for ref in (oms.backend_solve(rx, ry, rz), old.backend_solve(rx, ry, rz)):
            for a, b in zip(lg, ref):
                if len(b):
                    assert _rel(a, b) < 1e-9
        x, y, z = rng.standard_normal(n), rng.standard_normal(p), rng.standard_normal(m)
        assert _rel(be.eval_P_x(0.7, x), oms.backend_eval_P_x(0.7, x)) < 1e-12
        for a, b in zip(be.eval_A_xn_and_AT_xt(-1.0, 1.0, x, y), oms.backend_eval_A(-1.0, 1.0, x, y)):
            if len(b):
                assert _rel(a, b) < 1e-12
        for a, b in zip(be.eval_G_xn_and_GT_xt(1.0, 1.0, x, z), oms.backend_eval_G(1.0, 1.0, x, z)):
            if len(b):
                assert _rel(a, b) < 1e-12
    cl = be.clone()
    for a, b in zip(cl.solve(rx, ry, rz), be.solve(rx, ry, rz)):
        assert np.array_equal(a, b)


def test_reference_style_solver_drives_cuda_multistage(oracle, b200):
    """oracle KKTSystem + IP loop -> b200kkt_multistage_* through the C-ABI table: same iterations and solution"""
    q, g = load_scenario_mpc()
    st = oracle.default_settings(kkt_solver="sparse_multistage")
    cpu = oracle.SparseSolver(st); cpu.setup(*setup_args(q)); assert cpu.solve() == 1
    gpu = oracle.SparseSolver(st, backend_vtable=_vtable(oracle, b200)); gpu.setup(*setup_args(q)); assert gpu.solve() == 1
    rc, rg = cpu.result(), gpu.result()
    assert rg.info.iter == rc.info.iter == g["iterations"]
    assert np.abs(rg.x - rc.x).max() <= 1e-8 * max(1.0, np.abs(rc.x).max())


def test_batched_notebook_golden_trace(oracle, b200):
    """the device-resident IP loop + multistage CUDA backend reproduce the trace the REAL reference printed"""
    q, g = load_scenario_mpc()
    s = b200.SparseSolverBatched()
    s.settings.verbose = 2
    s.setup(3, q["P"], q["c"], q["A"], q["b"], None, None, None, q["x_l"], q["x_u"])
    assert [(d, o) for (_, d, o) in s.block_info()[:-1]] == [tuple(v) for v in g["multistage_block_sizes"]]
    infos = s.solve()
    assert [i.status for i in infos] == [1, 1, 1] and [i.iter for i in infos] == [12, 12, 12]
    assert abs(infos[0].primal_obj - g["objective"]) < 5e-2
    mine = trace_as_printed(s.trace(0)); golden = np.array(g["trace_sparse_multistage"])[:, 1:]
    assert mine.shape == golden.shape
    k = 7
    assert np.all(np.abs(mine[:k, :5] - golden[:k, :5]) <= 3e-5 * np.abs(golden[:k, :5]))
    assert np.all(np.abs(mine[:, 5:8] - golden[:, 5:8]) <= 6e-4 * np.abs(golden[:, 5:8]))
    assert np.abs(mine[:, 8:] - golden[:, 8:]).max() <= 6e-5
    r = s.result()
    assert np.array_equal(r.x[0], r.x[1]) and np.array_equal(r.x[0], r.x[2])     # identical instances -> bitwise identical results


def test_batched_mpc_matches_oracle(oracle, b200):
    """BASELINE config 4 shape at a short horizon and small batch: per-instance dynamics, same status / iterations / x as the oracle"""
    B = 6
    d = mpc_batch(B, N=25)
    s = b200.SparseSolverBatched()
    s.setup(B, d["P"], d["c"], d["A"], d["b"], None, None, None, d["x_l"], d["x_u"], Ax=d["Ax"])
    infos = s.solve(); r = s.result()
    for k in range(B):
        A = sp.csc_matrix((d["Ax"][k], d["A"].indices, d["A"].indptr), shape=d["A"].shape)
        o = oracle.SparseSolver(oracle.default_settings(kkt_solver="sparse_multistage")); o.setup(d["P"], d["c"][k], A, d["b"][k], None, None, None, d["x_l"][k], d["x_u"][k])
        st = o.solve(); ro = o.result()
        assert infos[k].status == st == 1
        assert infos[k].iter == ro.info.iter, (k, infos[k].iter, ro.info.iter)
        assert np.abs(r.x[k] - ro.x).max() <= 1e-8 * max(1.0, np.abs(ro.x).max())
        assert np.abs(r.y[k] - ro.y).max() <= 1e-5 * max(1.0, np.abs(ro.y).max())


@pytest.mark.parametrize("piece", ["Px", "Ax"])
def test_batched_partial_matrix_update_equals_fresh_setup(b200, piece):
    """update() of ONE matrix (Ruiz recomputed): the P and A^T A blocks of the multistage backend must both be refreshed"""
    B = 3
    d = mpc_batch(B, N=12)
    Px0 = np.broadcast_to(d["P"].data, (B, d["P"].nnz)).copy()
    def make(Px, Ax):
        s = b200.SparseSolverBatched()
        s.setup(B, d["P"], d["c"], d["A"], d["b"], None, None, None, d["x_l"], d["x_u"], Px=Px, Ax=Ax)
        return s
    s = make(Px0, d["Ax"]); s.solve()
    rng = np.random.default_rng(9)
    Px1, Ax1 = Px0, d["Ax"]
    if piece == "Px":
        Px1 = Px0 * rng.uniform(2.0, 5.0, Px0.shape)          # P is diagonal for the MPC problems: stays PSD
        s.update(Px=Px1)
    else:
        Ax1 = d["Ax"] * rng.uniform(0.7, 1.4, d["Ax"].shape)
        s.update(Ax=Ax1)
    iu = s.solve(); ru = s.result()
    f = make(Px1, Ax1); i_f = f.solve(); rf = f.result()
    for b in range(B):
        assert iu[b].status == i_f[b].status and iu[b].iter == i_f[b].iter, (b, iu[b].status, i_f[b].status, iu[b].iter, i_f[b].iter)
        if i_f[b].status == 1:
            assert np.abs(ru.x[b] - rf.x[b]).max() <= 1e-8 * max(1.0, np.abs(rf.x[b]).max())


def test_batched_sparse_known_answers_and_update(oracle, b200):
    """sparse/solver_test.cpp:67-107 golden values through the batched sparse API + the update() path"""
    q1 = simple_qp(); q2 = simple_qp_update(q1)
    S = lambda M: sp.csc_matrix(M)
    s = b200.SparseSolverBatched()
    # pattern from the union of both QPs' structures (identical here); instance 1 gets the second QP's values
    P1, A1, G1 = S(q1["P"]), S(q1["A"]), S(q1["G"])
    P2, A2 = S(q2["P"]), S(q2["A"])
    assert np.array_equal(P1.indices, P2.indices) and np.array_equal(A1.indices, A2.indices)
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
