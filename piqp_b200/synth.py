"""Synthetic QP generators with the distributions of the reference's test generator.

Follows include/piqp/utils/random_utils.hpp:131-286 (`rand::dense_strongly_convex_qp`,
`rand::sparse_strongly_convex_qp`): the DISTRIBUTIONS are reproduced, not libstdc++'s mt19937
stream (implementation-defined).  Instance b of a batch uses seed 42 + b.
"""
import numpy as np

INF = np.inf


def _bounds(rng, x_sol, G_x, n_ineq, dim, bounds_perc):
    delta_u = np.zeros(n_ineq)
    delta_l = np.zeros(n_ineq)
    for i in range(n_ineq):          # 30 % of the inequality constraints are inactive
        if rng.uniform() < 0.3:
            delta_u[i] = rng.uniform()
        if rng.uniform() < 0.3:
            delta_l[i] = rng.uniform()
    h_l = G_x - delta_l
    h_u = G_x + delta_u
    r = rng.uniform(size=n_ineq)
    h_l = np.where(r < 0.33, -INF, h_l)                      # 33 % upper only
    h_u = np.where((r >= 0.33) & (r < 0.66), INF, h_u)       # 33 % lower only
    x_l = np.full(dim, -INF)
    x_u = np.full(dim, INF)
    for i in range(dim):
        r = rng.uniform()
        if r < bounds_perc / 3:
            x_l[i] = x_sol[i]
            if rng.uniform() < 0.5:
                x_l[i] -= rng.uniform()
        elif r < bounds_perc * 2 / 3:
            x_u[i] = x_sol[i]
            if rng.uniform() < 0.5:
                x_u[i] += rng.uniform()
        elif r < bounds_perc:
            x_l[i] = x_sol[i]
            x_u[i] = x_sol[i]
            if rng.uniform() < 0.5:
                x_l[i] -= rng.uniform()
            else:
                x_u[i] += rng.uniform()
    return h_l, h_u, x_l, x_u


def dense_strongly_convex_qp(dim, n_eq, n_ineq, bounds_perc=0.5, strong_convexity_factor=1e-2, seed=42):
    """random_utils.hpp:131-208.  Returns dict(P (upper triangular), c, A, b, G, h_l, h_u, x_l, x_u)."""
    rng = np.random.default_rng(seed)
    P = np.triu(rng.standard_normal((dim, dim)), 1)
    lam_min = np.linalg.eigvalsh(P + P.T - np.diag(np.diag(P))).min() if dim > 0 else 0.0
    P = P + np.eye(dim) * (strong_convexity_factor + abs(lam_min))
    A = rng.standard_normal((n_eq, dim))
    G = rng.standard_normal((n_ineq, dim))
    x_sol = rng.standard_normal(dim)
    c = rng.standard_normal(dim)
    b = A @ x_sol
    h_l, h_u, x_l, x_u = _bounds(rng, x_sol, G @ x_sol, n_ineq, dim, bounds_perc)
    return dict(P=P, c=c, A=A, b=b, G=G, h_l=h_l, h_u=h_u, x_l=x_l, x_u=x_u, x_sol=x_sol)


def sparse_strongly_convex_qp(dim, n_eq, n_ineq, sparsity_factor, bounds_perc=0.5, strong_convexity_factor=1e-2, seed=42,
                              eig_shift=None):
    """random_utils.hpp:211-286 with scipy.sparse matrices (P upper triangular CSC).

    For large dim the reference's dense eigenvalue solve is replaced by a Gershgorin bound
    (eig_shift='gershgorin') so the generator stays O(nnz); the matrix is still strictly PD.
    """
    import scipy.sparse as sp
    rng = np.random.default_rng(seed)

    def rand_sparse(r, c, upper=False):
        nnz = rng.binomial(r * c, sparsity_factor) if r * c > 0 else 0
        ri = rng.integers(0, max(r, 1), nnz)
        ci = rng.integers(0, max(c, 1), nnz)
        M = sp.coo_matrix((rng.standard_normal(nnz), (ri, ci)), shape=(r, c)).tocsc()
        M.sum_duplicates()
        return sp.triu(M, 1, format="csc") if upper else M

    P = rand_sparse(dim, dim, upper=True)
    if eig_shift is None:
        eig_shift = "exact" if dim <= 2000 else "gershgorin"
    if eig_shift == "exact":
        lam_min = np.linalg.eigvalsh((P + P.T).toarray()).min() if dim > 0 else 0.0
    else:
        S = abs(P) + abs(P.T)
        lam_min = -np.asarray(S.sum(axis=1)).ravel().max() if dim > 0 else 0.0
    P = (P + sp.identity(dim, format="csc") * (strong_convexity_factor + abs(lam_min))).tocsc()
    A = rand_sparse(n_eq, dim)
    G = rand_sparse(n_ineq, dim)
    x_sol = rng.standard_normal(dim)
    c = rng.standard_normal(dim)
    b = A @ x_sol
    h_l, h_u, x_l, x_u = _bounds(rng, x_sol, G @ x_sol, n_ineq, dim, bounds_perc)
    return dict(P=P, c=c, A=A, b=b, G=G, h_l=h_l, h_u=h_u, x_l=x_l, x_u=x_u, x_sol=x_sol)
