"""GPU: dense LDL^T without pivoting (piqp::dense::LDLTNoPivot, include/piqp/dense/ldlt_no_pivot.hpp) through the
sparse multifrontal backend's single-front path.  Mirrors tests/src/dense/ldlt_test.cpp (SolveLower / SolveUpper: b ~ P x to 1e-8)
and adds the oracle's restatement of the blocked algorithm, a quasi-definite (indefinite) matrix, refactorisation and a size
that takes the whole-GPU blocked path (n >= 512: 64-column panels + DMMA trailing updates + blocked solves)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _pd(n, seed):
    rng = np.random.default_rng(seed)
    M = rng.standard_normal((n, n))
    return M @ M.T / n + np.eye(n)


@pytest.mark.parametrize("uplo", ["lower", "upper"])
@pytest.mark.parametrize("dim", [50, 300, 700])
def test_dense_ldlt_solve(oracle, b200, dim, uplo):
    P = _pd(dim, dim)
    tri = np.tril(P) if uplo == "lower" else np.triu(P)
    ldlt = b200.LDLTNoPivot()
    assert ldlt.compute(tri, uplo).info() is True
    assert ldlt.compute(tri, uplo).info() is True            # recompute on the same handle (the reference checks: no allocation)
    b = np.random.default_rng(1).standard_normal(dim)
    x = ldlt.solve(b)
    assert np.linalg.norm(P @ x - b) <= 1e-8 * min(np.linalg.norm(b), np.linalg.norm(P @ x))      # Eigen isApprox(1e-8)
    fac = oracle.ldlt(P)                                     # (unit L, D, info, factored matrix)
    xo = oracle.ldlt_solve(fac[3], b)
    assert np.abs(x - xo).max() <= 1e-9 * max(1.0, np.abs(xo).max())


@pytest.mark.parametrize("n1,n2", [(40, 25), (400, 250)])
def test_dense_ldlt_quasi_definite(b200, n1, n2):
    """[[H, B^T], [B, -C]] with H, C positive definite: negative pivots, still no pivoting needed"""
    rng = np.random.default_rng(7)
    H, Cm, B = _pd(n1, 3), _pd(n2, 4), rng.standard_normal((n2, n1))
    K = np.block([[H, B.T], [B, -Cm]])
    ldlt = b200.LDLTNoPivot()
    assert ldlt.compute(np.tril(K)).info() is True
    b = rng.standard_normal(n1 + n2)
    x = ldlt.solve(b)
    assert np.abs(K @ x - b).max() <= 1e-9 * max(1.0, np.abs(x).max())
    K2 = K.copy(); K2[:n1, :n1] += 0.5 * np.eye(n1)
    assert ldlt.compute(np.tril(K2)).info() is True           # new values, same handle
    x2 = ldlt.solve(b)
    assert np.abs(K2 @ x2 - b).max() <= 1e-9 * max(1.0, np.abs(x2).max())


def test_dense_ldlt_reports_zero_pivot(b200):
    K = np.zeros((6, 6)); K[np.arange(6), np.arange(6)] = [1, 2, 0, 4, 5, 6]
    assert b200.LDLTNoPivot().compute(K).info() is False      # ldlt_no_pivot.hpp: info() != Success
