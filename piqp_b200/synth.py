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


def dense_batch_torch(batch, n, p, m, seed0=42, device="cuda", bounds_perc=0.5, strong_convexity_factor=1e-2):
    """Batched, device-side version of dense_strongly_convex_qp (same distributions, torch's Philox stream):
    instance b is drawn from a generator seeded with seed0 + b.  Returns a dict of float64 torch tensors
    (P [B,n,n] upper triangular, c, A, b, G, h_l, h_u, x_l, x_u) resident on `device`."""
    import torch
    dd = dict(dtype=torch.float64, device=device)
    out = {k: [] for k in ("P", "c", "A", "b", "G", "h_l", "h_u", "x_l", "x_u")}
    inf = float("inf")
    for i in range(batch):
        g = torch.Generator(device=device)
        g.manual_seed(seed0 + i)
        rn = lambda *s: torch.randn(*s, generator=g, **dd)
        ru = lambda *s: torch.rand(*s, generator=g, **dd)
        P = torch.triu(rn(n, n), 1)
        lam_min = torch.linalg.eigvalsh(P + P.T).min()
        P = P + torch.eye(n, **dd) * (strong_convexity_factor + lam_min.abs())
        A = rn(p, n); G = rn(m, n); x_sol = rn(n); c = rn(n)
        b = A @ x_sol
        delta_u = torch.where(ru(m) < 0.3, ru(m), torch.zeros(m, **dd))
        delta_l = torch.where(ru(m) < 0.3, ru(m), torch.zeros(m, **dd))
        Gx = G @ x_sol
        h_l = Gx - delta_l; h_u = Gx + delta_u
        r = ru(m)
        h_l = torch.where(r < 0.33, torch.full_like(h_l, -inf), h_l)
        h_u = torch.where((r >= 0.33) & (r < 0.66), torch.full_like(h_u, inf), h_u)
        r = ru(n); act = ru(n) < 0.5; shift = ru(n); side = ru(n) < 0.5
        lo = r < bounds_perc / 3
        up = (r >= bounds_perc / 3) & (r < bounds_perc * 2 / 3)
        both = (r >= bounds_perc * 2 / 3) & (r < bounds_perc)
        x_l = torch.full((n,), -inf, **dd); x_u = torch.full((n,), inf, **dd)
        x_l = torch.where(lo, x_sol - torch.where(act, shift, torch.zeros_like(shift)), x_l)
        x_u = torch.where(up, x_sol + torch.where(act, shift, torch.zeros_like(shift)), x_u)
        x_l = torch.where(both, x_sol - torch.where(side, shift, torch.zeros_like(shift)), x_l)
        x_u = torch.where(both, x_sol + torch.where(~side, shift, torch.zeros_like(shift)), x_u)
        for k, v in (("P", P), ("c", c), ("A", A), ("b", b), ("G", G), ("h_l", h_l), ("h_u", h_u), ("x_l", x_l), ("x_u", x_u)):
            out[k].append(v)
    return {k: torch.stack(v).contiguous() for k, v in out.items()}


def mpc_batch(batch, N=100, nx=12, nu=4, seed0=42, umax=1.0, xmax=5.0):
    """BASELINE config 4: linear time-invariant MPC QPs in the variable order the multistage backend requires
    (docs/_pages/multistage.md:72): z = (x_0, u_0, x_1, u_1, ..., x_{N-1}, u_{N-1}, x_N), n = N (nx + nu) + nx,
    dynamics equalities x_{i+1} = A_d x_i + B_d u_i (p = N nx), x_0 fixed through equal box bounds, box bounds on
    states and inputs (handled in x_reg, no G), stage cost Q = I, R = 0.1 I, terminal cost 10 I.
    Every instance b has its own random stable (A_d, B_d) and x_0 (seed0 + b); all share the sparsity pattern.
    Returns dict(P, A: scipy CSC patterns (values of instance 0), Ax [batch, nnz(A)], c, b, x_l, x_u [batch, *])."""
    import scipy.sparse as sp
    nz = nx + nu
    n = N * nz + nx
    p = N * nx
    Pd = np.concatenate([np.tile(np.concatenate([np.ones(nx), 0.1 * np.ones(nu)]), N), 10.0 * np.ones(nx)])
    P = sp.diags(Pd).tocsc()
    rows, cols = [], []
    for i in range(N):
        r0 = i * nx
        for r in range(nx):
            for cc in range(nz):                      # [A_d B_d] on (x_i, u_i)
                rows.append(r0 + r); cols.append(i * nz + cc)
            rows.append(r0 + r); cols.append((i + 1) * nz + r)   # -I on x_{i+1}
    rows = np.array(rows); cols = np.array(cols)
    A_pat = sp.csc_matrix((np.ones(len(rows)), (rows, cols)), shape=(p, n))
    A_pat.sort_indices()
    # position of every (row, col) in the CSC value array
    order = np.lexsort((rows, cols))
    inv = np.empty_like(order); inv[order] = np.arange(len(order))
    Ax = np.zeros((batch, A_pat.nnz)); c = np.zeros((batch, n)); b = np.zeros((batch, p))
    x_l = np.full((batch, n), -np.inf); x_u = np.full((batch, n), np.inf)
    for k in range(batch):
        rng = np.random.default_rng(seed0 + k)
        Ad = rng.standard_normal((nx, nx)); Ad *= 0.95 / max(abs(np.linalg.eigvals(Ad)))
        Bd = rng.standard_normal((nx, nu)) / np.sqrt(nx)
        x0 = rng.uniform(-1.0, 1.0, nx)
        vals = np.zeros(len(rows)); t = 0
        blk = np.hstack([Ad, Bd])
        for i in range(N):
            for r in range(nx):
                vals[t:t + nz] = blk[r]; t += nz
                vals[t] = -1.0; t += 1
        Ax[k, inv] = vals
        for i in range(N):
            x_l[k, i * nz:i * nz + nx] = -xmax; x_u[k, i * nz:i * nz + nx] = xmax
            x_l[k, i * nz + nx:(i + 1) * nz] = -umax; x_u[k, i * nz + nx:(i + 1) * nz] = umax
        x_l[k, N * nz:] = -xmax; x_u[k, N * nz:] = xmax
        x_l[k, :nx] = x0; x_u[k, :nx] = x0
    A0 = sp.csc_matrix((Ax[0], A_pat.indices, A_pat.indptr), shape=(p, n))
    return dict(P=P, A=A0, Ax=Ax, c=c, b=b, x_l=x_l, x_u=x_u, n=n, p=p)


def sparse_batch(batch, n=1000, p=500, m=500, density=0.005, seed0=42):
    """BASELINE config 3 / 5 family: `batch` random sparse strongly convex QPs (random_utils.hpp:211-286) that share one
    sparsity pattern (the batched sparse API's contract) and differ in values: instance k scales diag(P) by U(1, 1.2),
    perturbs every entry of A and G by U(0.8, 1.2), draws its own c and x_sol, and gets b = A_k x_sol and bounds around
    G_k x_sol / x_sol by the reference's recipe (so every instance is feasible).
    Returns scipy patterns P (upper), A, G and arrays Px, Ax, Gx, c, b, h_l, h_u, x_l, x_u with a leading batch axis."""
    import scipy.sparse as sp
    base = sparse_strongly_convex_qp(n, p, m, density, seed=seed0, eig_shift="gershgorin")
    Pu = sp.csc_matrix(sp.triu(base["P"])); A = sp.csc_matrix(base["A"]); G = sp.csc_matrix(base["G"])
    Pu.sort_indices(); A.sort_indices(); G.sort_indices()
    diag_mask = np.repeat(np.arange(n), np.diff(Pu.indptr)) == Pu.indices
    out = {k: [] for k in ("Px", "Ax", "Gx", "c", "b", "h_l", "h_u", "x_l", "x_u")}
    for k in range(batch):
        rng = np.random.default_rng(seed0 + 7919 * (k + 1))
        Px = Pu.data * np.where(diag_mask, rng.uniform(1.0, 1.2), 1.0)
        Ax = A.data * rng.uniform(0.8, 1.2, A.nnz)
        Gx = G.data * rng.uniform(0.8, 1.2, G.nnz)
        Ak = sp.csc_matrix((Ax, A.indices, A.indptr), shape=A.shape); Gk = sp.csc_matrix((Gx, G.indices, G.indptr), shape=G.shape)
        x_sol = rng.standard_normal(n)
        h_l, h_u, x_l, x_u = _bounds(rng, x_sol, Gk @ x_sol, m, n, 0.5)
        for key, v in (("Px", Px), ("Ax", Ax), ("Gx", Gx), ("c", rng.standard_normal(n)), ("b", Ak @ x_sol), ("h_l", h_l), ("h_u", h_u), ("x_l", x_l), ("x_u", x_u)):
            out[key].append(v)
    out = {k: np.stack(v) for k, v in out.items()}
    out.update(P=Pu, A=A, G=G, n=n, p=p, m=m)
    return out


def kkt_min_degree_perm(P, A, G):
    """A fill-reducing permutation of the FULL KKT pattern [[P, A^T, G^T], [A, I, 0], [G, 0, I]] from SuperLU's
    multiple-minimum-degree code (scipy), standing in for Eigen::AMDOrdering (sparse/ordering.hpp:68-84) on the CPU arm."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    n, p, m = P.shape[0], A.shape[0], G.shape[0]
    Pf = sp.triu(P, 1); Pf = abs(Pf) + abs(Pf.T)
    K = sp.bmat([[Pf, abs(A.T), abs(G.T)], [abs(A), None, None], [abs(G), None, None]], format="csc") + sp.identity(n + p + m, format="csc") * (1.0 + 4.0 * (n + p + m))
    lu = spla.splu(K, permc_spec="MMD_AT_PLUS_A", diag_pivot_thresh=0.0, options=dict(SymmetricMode=True))
    return np.asarray(lu.perm_c, dtype=np.int32)
