"""CPU tests of the oracle against the reference's own known answers (no GPU needed).

Golden values: /root/reference/tests/src/dense/solver_test.cpp:60-100, 373-376; statuses :107-182.
"""
import numpy as np
import pytest

from helpers import (dual_infeasible_qp, ill_conditioned_qp, inf_bounds_qp, kkt_residuals, primal_infeasible_qp, setup_args,
                     simple_qp, simple_qp_update)
from piqp_b200.synth import dense_strongly_convex_qp


def test_simple_qp_with_update(oracle):
    s = oracle.DenseSolver()
    q = simple_qp()
    s.setup(*setup_args(q))
    assert s.solve() == 1
    r = s.result()
    assert abs(r.x[0] - 0.4285714) < 1e-6 and abs(r.x[1] - 0.2142857) < 1e-6
    assert abs(r.y[0] + 1.5714286) < 1e-6
    for v in (r.z_l, r.z_u, r.z_bl, r.z_bu):
        assert np.abs(v).max() < 1e-6
    q2 = simple_qp_update(q)
    s.update(P=q2["P"], c=q2["c"], A=q2["A"], b=q2["b"], h_u=q2["h_u"], x_u=q2["x_u"])
    assert s.solve() == 1
    r = s.result()
    assert abs(r.x[0] - 0.2763157) < 1e-6 and abs(r.x[1] - 0.0921056) < 1e-6
    assert abs(r.y[0] + 1.2105263) < 1e-6
    for v in (r.z_l, r.z_u, r.z_bl, r.z_bu):
        assert np.abs(v).max() < 1e-6


def test_infeasibility_statuses(oracle):
    s = oracle.DenseSolver(); s.setup(*setup_args(primal_infeasible_qp()))
    assert s.solve() == -2
    s = oracle.DenseSolver(); s.setup(*setup_args(dual_infeasible_qp()))
    assert s.solve() == -3


def test_ill_conditioned_and_inf_bounds(oracle):
    s = oracle.DenseSolver(); s.setup(*setup_args(ill_conditioned_qp()))
    assert s.solve() == 1
    s = oracle.DenseSolver(); s.setup(*setup_args(inf_bounds_qp()))
    assert s.solve() == 1
    assert np.allclose(s.result().x, [-0.5, -1.0, -0.5, -1.0], atol=1e-6)


@pytest.mark.parametrize("dims,kw", [((20, 10, 12), {}), ((20, 10, 12), dict(strong_convexity_factor=0.0)),
                                      ((64, 10, 0), dict(bounds_perc=0.0)), ((20, 0, 12), {}), ((64, 0, 0), dict(bounds_perc=0.0)),
                                      ((128, 32, 64), {})])
def test_random_qps_solve(oracle, dims, kw):
    """solver_test.cpp:206-345: every generated QP must reach PIQP_SOLVED; we also check the KKT conditions."""
    q = dense_strongly_convex_qp(*dims, seed=42, **kw)
    s = oracle.DenseSolver(); s.setup(*setup_args(q))
    assert s.solve() == 1
    assert kkt_residuals(q, s.result()) < 1e-6


def test_same_result_with_ruiz(oracle):
    """solver_test.cpp:244-288"""
    q = dense_strongly_convex_qp(20, 10, 12, strong_convexity_factor=0.0, seed=7)
    a = oracle.DenseSolver(oracle.default_settings(eps_rel=0.0), identity_preconditioner=True); a.setup(*setup_args(q))
    b = oracle.DenseSolver(oracle.default_settings(eps_rel=0.0)); b.setup(*setup_args(q))
    assert a.solve() == 1 and b.solve() == 1
    assert np.linalg.norm(a.result().x - b.result().x) < 1e-6


def test_cholesky_and_ldlt_against_numpy(oracle):
    """tests/src/dense/ldlt_test.cpp:22-77 (b = P x to 1e-8) for both factorisations, several block regimes"""
    rng = np.random.default_rng(0)
    for n in (5, 31, 50, 200, 300):
        M = rng.standard_normal((n, n)); S = M @ M.T + n * np.eye(n)
        L, info = oracle.chol(S)
        assert info == -1
        assert np.allclose(L @ L.T, S, rtol=1e-12, atol=1e-10)
        assert np.allclose(L, np.linalg.cholesky(S), rtol=1e-10, atol=1e-10)
        Lu, D, info, fac = oracle.ldlt(S)
        assert info == -1
        assert np.allclose(Lu @ np.diag(D) @ Lu.T, S, rtol=1e-12, atol=1e-9)
        b = rng.standard_normal(n)
        x = oracle.ldlt_solve(fac, b)
        assert np.allclose(S @ x, b, rtol=1e-8, atol=1e-8)
    Lf, info = oracle.chol(np.array([[1., 2], [2, 1]]))
    assert info == 1   # not positive definite: fails at column 1


def test_kktsystem_factorize_solve_roundtrip(oracle):
    """tests/src/dense/kkt_test.cpp:67-139: KKTSystem.solve then KKTSystem.mul reproduces the rhs to 1e-8"""
    dim, n_eq, n_ineq = 20, 8, 9
    q = dense_strongly_convex_qp(dim, n_eq, n_ineq, seed=3)
    s = oracle.DenseSolver(identity_preconditioner=True); s.setup(*setup_args(q))
    n, p, m, nhl, nhu, nxl, nxu = s.dims
    N = 5 * n + p + 4 * m
    rng = np.random.default_rng(1)
    scaling = np.ones(N); rhs = rng.standard_normal(N)
    for ir in (False, True):
        ok, lhs, back = s.kktsystem_roundtrip(0.9, 1.2, scaling, rhs, iterative_refinement=ir)
        assert ok == 1
        o = 0
        blocks = {}
        for name, k in (("x", n), ("y", p), ("z_l", m), ("z_u", m), ("z_bl", n), ("z_bu", n), ("s_l", m), ("s_u", m), ("s_bl", n), ("s_bu", n)):
            blocks[name] = (rhs[o:o + k], back[o:o + k]); o += k
        tol = 1e-8 if not ir else 1e-6   # with refinement the static regularisation perturbs the operator (kkt_system.hpp:205)
        assert np.allclose(*blocks["x"], atol=tol) and np.allclose(*blocks["y"], atol=tol)
        hl = np.isfinite(q["h_l"]); hu = np.isfinite(q["h_u"])
        assert np.allclose(blocks["z_l"][0][hl], blocks["z_l"][1][hl], atol=tol) and np.allclose(blocks["s_l"][0][hl], blocks["s_l"][1][hl], atol=tol)
        assert np.allclose(blocks["z_u"][0][hu], blocks["z_u"][1][hu], atol=tol) and np.allclose(blocks["s_u"][0][hu], blocks["s_u"][1][hu], atol=tol)
        assert np.allclose(blocks["z_bl"][0][:nxl], blocks["z_bl"][1][:nxl], atol=tol) and np.allclose(blocks["z_bu"][0][:nxu], blocks["z_bu"][1][:nxu], atol=tol)
        assert np.allclose(blocks["s_bl"][0][:nxl], blocks["s_bl"][1][:nxl], atol=tol) and np.allclose(blocks["s_bu"][0][:nxu], blocks["s_bu"][1][:nxu], atol=tol)


def test_dense_kkt_assembly_matches_numpy(oracle):
    """dense::KKT::update_kkt (dense/kkt.hpp:140-160) against a NumPy evaluation of P + diag + AtA/delta + G^T Z^-1 G"""
    q = dense_strongly_convex_qp(30, 7, 11, seed=5)
    s = oracle.DenseSolver(); s.setup(*setup_args(q))
    P, AT, GT = s.scaled_matrices()
    rng = np.random.default_rng(2)
    x_reg = rng.uniform(0.1, 1, 30); z_reg = rng.uniform(0.1, 1, 11); delta = 0.37
    assert s.backend_factor(delta, x_reg, z_reg) == 1
    K, L = s.kkt_and_factor()
    Pf = P + P.T - np.diag(np.diag(P))
    ref = Pf + np.diag(x_reg) + AT @ AT.T / delta + GT @ np.diag(1 / z_reg) @ GT.T
    assert np.allclose(np.tril(K), np.tril(ref), rtol=1e-12, atol=1e-12)
    assert np.allclose(L @ L.T, ref, rtol=1e-10, atol=1e-10)
    rx, ry, rz = rng.standard_normal(30), rng.standard_normal(7), rng.standard_normal(11)
    lx, ly, lz = s.backend_solve(rx, ry, rz)
    # the 3x3 system the backend solves (kkt_solver_base.hpp:34)
    assert np.allclose(Pf @ lx + x_reg * lx + AT @ ly + GT @ lz, rx, atol=1e-9)
    assert np.allclose(AT.T @ lx - delta * ly, ry, atol=1e-9)
    assert np.allclose(GT.T @ lx - z_reg * lz, rz, atol=1e-9)


def test_ruiz_equilibrates(oracle):
    """after Ruiz the scaled KKT columns have inf-norm close to 1 (dense/preconditioner.hpp:64-165)"""
    q = dense_strongly_convex_qp(40, 10, 20, seed=9)
    q["P"] = q["P"] * 100; q["G"] = q["G"] * 1e-2
    s = oracle.DenseSolver(); s.setup(*setup_args(q))
    P, AT, GT = s.scaled_matrices()
    Pf = P + P.T - np.diag(np.diag(P))
    Kfull = np.block([[Pf, AT, GT], [AT.T, np.zeros((10, 10)), np.zeros((10, 20))], [GT.T, np.zeros((20, 10)), np.zeros((20, 20))]])
    norms = np.abs(Kfull).max(axis=0)
    assert norms.max() < 1.6 and norms.min() > 0.3
