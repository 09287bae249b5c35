"""Regenerates the scenario-MPC QP of the reference's documentation notebook
(/root/reference/docs/assets/robust_scenario_mpc.ipynb, cells 1-23: np.random.seed(42), M=3, N=5, Ns=3) and stores
it together with the solver output the notebook prints (cells 25 and 27: per-iteration trace of sparse_ldlt and
sparse_multistage, detected block sizes, arrow width).  Those printed numbers were produced by the REAL reference
and are the golden vectors that pin the oracle's IP loop and the multistage structure detection.

Run in the build container:  python tests/golden/make_scenario_mpc.py
"""
import json
import os

import numpy as np
import scipy.sparse as sp
from scipy.linalg import solve_discrete_are
from scipy.signal import cont2discrete

HERE = os.path.dirname(os.path.abspath(__file__))


class ChainMassSystem:
    def __init__(self, M, m=1.0, c=0.1, k=1.0):
        self.M, self.m, self.c, self.k = M, m, c, k
        self.nx_max, self.nu_max = 4.0, 0.5
        self.nx, self.nu = 2 * M, M - 1
        L = np.eye(M, k=-1)
        A = np.block([[np.zeros((M, M)), np.eye(M)], [(-2 * k * np.eye(M) + k * L + k * L.T) / m, (-2 * c * np.eye(M)) / m]])
        B = np.block([[np.zeros((2 * M - self.nu, self.nu))], [np.eye(self.nu)]])
        d = cont2discrete((A, B, np.eye(self.nx), np.zeros((self.nx, self.nu))), 0.5, method="zoh")
        self.Ad, self.Bd = d[0], d[1]
        self.Q = 1e3 * np.eye(self.nx)
        self.R = 1e-1 * np.eye(self.nu)
        self.QN = solve_discrete_are(self.Ad, self.Bd, self.Q, self.R)


def build(M, N, Ns, rng_seed=42):
    np.random.seed(rng_seed)
    systems = [ChainMassSystem(M, k=k) for k in np.linspace(1.0, 2.0, Ns)]
    nx, nu = systems[0].nx, systems[0].nu
    x0 = np.random.uniform(-1.0, 1.0, nx)
    n = nx + nu + Ns * ((N - 1) * (nx + nu) + nx)
    p = Ns * N * nx
    # the notebook assigns dense blocks into csc matrices, which stores the blocks' zeros explicitly;
    # the sparsity STRUCTURE (incl. those zeros) is what the reference's block detection sees, so keep it
    class Blocks:
        def __init__(self, shape):
            self.shape, self.r, self.c, self.v = shape, [], [], []
        def __setitem__(self, key, val):
            rs, cs = key
            val = np.asarray(val, dtype=float)
            r0 = rs.start + (self.shape[0] if rs.start < 0 else 0); r1 = (rs.stop if rs.stop is not None else self.shape[0]); r1 += self.shape[0] if r1 <= 0 and rs.stop is not None and rs.stop < 0 else 0
            c0 = cs.start + (self.shape[1] if cs.start < 0 else 0); c1 = (cs.stop if cs.stop is not None else self.shape[1]); c1 += self.shape[1] if c1 <= 0 and cs.stop is not None and cs.stop < 0 else 0
            assert val.shape == (r1 - r0, c1 - c0), (val.shape, r0, r1, c0, c1)
            rr, cc = np.meshgrid(np.arange(r0, r1), np.arange(c0, c1), indexing="ij")
            self.r += rr.ravel().tolist(); self.c += cc.ravel().tolist(); self.v += val.ravel().tolist()
        def tocsc(self):
            M = sp.coo_matrix((self.v, (self.r, self.c)), shape=self.shape).tocsc()
            M.sort_indices()
            return M
    P = Blocks((n, n)); A = Blocks((p, n))
    c = np.zeros(n); b = np.zeros(p); x_l = np.zeros(n); x_u = np.zeros(n)
    x_l[-(nx + nu):-nu] = x0; x_u[-(nx + nu):-nu] = x0
    x_l[-nu:] = -systems[0].nu_max; x_u[-nu:] = systems[0].nu_max
    P[-(nx + nu):-nu, -(nx + nu):-nu] = systems[0].Q
    P[-nu:, -nu:] = systems[0].R
    for s in range(Ns):
        off = s * ((N - 1) * (nx + nu) + nx) - (nx + nu)
        for i in range(1, N):
            P[off + i * (nx + nu):off + i * (nx + nu) + nx, off + i * (nx + nu):off + i * (nx + nu) + nx] = systems[s].Q / Ns
            P[off + i * (nx + nu) + nx:off + i * (nx + nu) + nx + nu, off + i * (nx + nu) + nx:off + i * (nx + nu) + nx + nu] = systems[s].R / Ns
        P[off + N * (nx + nu):off + N * (nx + nu) + nx, off + N * (nx + nu):off + N * (nx + nu) + nx] = systems[s].QN / Ns
    for s in range(Ns):
        off = s * ((N - 1) * (nx + nu) + nx) - (nx + nu)
        eo = s * N * nx
        for i in range(N):
            if i == 0:
                A[eo + i * nx:eo + (i + 1) * nx, -(nx + nu):-nu] = systems[s].Ad
                A[eo + i * nx:eo + (i + 1) * nx, -nu:] = systems[s].Bd
            else:
                A[eo + i * nx:eo + (i + 1) * nx, off + i * (nx + nu):off + i * (nx + nu) + nx] = systems[s].Ad
                A[eo + i * nx:eo + (i + 1) * nx, off + i * (nx + nu) + nx:off + i * (nx + nu) + nx + nu] = systems[s].Bd
            A[eo + i * nx:eo + (i + 1) * nx, off + (i + 1) * (nx + nu):off + (i + 1) * (nx + nu) + nx] = -np.eye(nx)
    for s in range(Ns):
        off = s * ((N - 1) * (nx + nu) + nx) - (nx + nu)
        for i in range(N):
            if i > 0:
                x_l[off + i * (nx + nu) + nx:off + i * (nx + nu) + nx + nu] = -systems[s].nu_max
                x_u[off + i * (nx + nu) + nx:off + i * (nx + nu) + nx + nu] = systems[s].nu_max
            x_l[off + (i + 1) * (nx + nu):off + (i + 1) * (nx + nu) + nx] = -systems[s].nx_max
            x_u[off + (i + 1) * (nx + nu):off + (i + 1) * (nx + nu) + nx] = systems[s].nx_max
    return P.tocsc(), c, A.tocsc(), b, x_l, x_u


# iter prim_obj dual_obj duality_gap prim_res dual_res rho delta mu p_step d_step  -- notebook cell 25 (sparse_ldlt)
TRACE_LDLT = """
0 3.25459e+02 -1.09791e+06 1.09824e+06 1.93609e-03 6.63672e+02 1.000e-06 1.000e-04 1.177e+04 0.0000 0.0000
1 7.50453e+02 -2.77013e+05 2.77764e+05 1.83182e-03 2.15106e+01 1.450e-07 1.450e-05 1.706e+03 0.8673 0.9900
2 3.18009e+03 -2.17709e+04 2.49510e+04 8.92483e-04 2.82825e+01 5.695e-08 1.278e-06 1.504e+02 0.8896 0.9398
3 3.56810e+03 2.22331e+03 1.34479e+03 3.99874e-04 1.98075e+02 2.444e-08 1.825e-07 2.148e+01 0.6191 0.9808
4 4.22500e+03 4.04344e+03 1.81561e+02 9.04618e-05 1.01304e+02 5.476e-09 4.090e-08 4.813e+00 0.7965 0.9681
5 4.46048e+03 4.38640e+03 7.40831e+01 1.69913e-06 4.08054e+00 4.367e-10 3.262e-09 3.838e-01 0.9784 0.9075
6 4.45291e+03 4.44882e+03 4.09229e+00 4.39455e-08 8.67742e-02 1.000e-10 1.625e-10 1.912e-02 0.9678 0.9702
7 4.45183e+03 4.45162e+03 2.14879e-01 8.47502e-10 1.09070e-01 1.000e-10 1.000e-10 9.685e-04 0.9793 0.9697
8 4.45174e+03 4.45173e+03 1.19518e-02 8.97259e-11 2.23268e-02 1.000e-10 1.000e-10 5.190e-05 0.9884 0.9834
9 4.45173e+03 4.45173e+03 1.33397e-03 3.31985e-11 2.23268e-04 1.000e-10 1.000e-10 5.501e-06 0.9900 0.9900
10 4.45173e+03 4.45173e+03 1.95468e-04 3.10567e-11 2.23263e-06 1.000e-10 1.000e-10 8.052e-07 0.9900 0.9900
11 4.45173e+03 4.45173e+03 2.80912e-05 1.25559e-11 2.23069e-08 1.000e-10 1.000e-10 1.157e-07 0.9900 0.9900
12 4.45173e+03 4.45173e+03 3.83423e-06 4.68808e-12 2.15834e-10 1.000e-10 1.000e-10 1.578e-08 0.9900 0.9900
"""
# notebook cell 27 (sparse_multistage)
TRACE_MULTISTAGE = """
0 3.25459e+02 -1.09791e+06 1.09824e+06 1.93609e-03 6.63672e+02 1.000e-06 1.000e-04 1.177e+04 0.0000 0.0000
1 7.50453e+02 -2.77013e+05 2.77764e+05 1.83182e-03 2.15106e+01 1.450e-07 1.450e-05 1.706e+03 0.8673 0.9900
2 3.18009e+03 -2.17709e+04 2.49510e+04 8.92483e-04 2.82825e+01 5.695e-08 1.278e-06 1.504e+02 0.8896 0.9398
3 3.56810e+03 2.22331e+03 1.34479e+03 3.99874e-04 1.98075e+02 2.444e-08 1.825e-07 2.148e+01 0.6191 0.9808
4 4.22500e+03 4.04344e+03 1.81561e+02 9.04618e-05 1.01304e+02 5.476e-09 4.090e-08 4.813e+00 0.7965 0.9681
5 4.46048e+03 4.38640e+03 7.40831e+01 1.69913e-06 4.08054e+00 4.367e-10 3.262e-09 3.838e-01 0.9784 0.9075
6 4.45291e+03 4.44882e+03 4.09229e+00 4.39455e-08 8.67741e-02 1.000e-10 1.625e-10 1.912e-02 0.9678 0.9702
7 4.45183e+03 4.45162e+03 2.14879e-01 8.47502e-10 1.09069e-01 1.000e-10 1.000e-10 9.685e-04 0.9793 0.9697
8 4.45174e+03 4.45173e+03 1.19519e-02 8.97259e-11 2.23266e-02 1.000e-10 1.000e-10 5.190e-05 0.9884 0.9834
9 4.45173e+03 4.45173e+03 1.33393e-03 3.31984e-11 2.23276e-04 1.000e-10 1.000e-10 5.501e-06 0.9900 0.9900
10 4.45173e+03 4.45173e+03 1.95466e-04 3.10567e-11 2.16650e-06 1.000e-10 1.000e-10 8.052e-07 0.9900 0.9900
11 4.45173e+03 4.45173e+03 2.80900e-05 1.25559e-11 3.04364e-08 1.000e-10 1.000e-10 1.157e-07 0.9900 0.9900
12 4.45173e+03 4.45173e+03 3.83881e-06 4.68814e-12 1.54665e-08 1.000e-10 1.000e-10 1.578e-08 0.9900 0.9900
"""
GOLDEN = {
    "source": "docs/assets/robust_scenario_mpc.ipynb:489-573 (printed output of the real reference)",
    "n": 122, "p": 90, "nnz_P_utri": 375, "nnz_A": 1260, "iterations": 12, "objective": 4.45173e+03,
    "multistage_block_sizes": [[8, 6], [8, 6], [8, 6], [14, 0]] * 3, "multistage_arrow_width": 8,
    "trace_columns": ["iter", "prim_obj", "dual_obj", "duality_gap", "prim_res", "dual_res", "rho", "delta", "mu", "p_step", "d_step"],
    "trace_sparse_ldlt": [[float(v) for v in line.split()] for line in TRACE_LDLT.strip().splitlines()],
    "trace_sparse_multistage": [[float(v) for v in line.split()] for line in TRACE_MULTISTAGE.strip().splitlines()],
}

if __name__ == "__main__":
    P, c, A, b, x_l, x_u = build(3, 5, 3)
    assert P.shape == (122, 122) and A.shape == (90, 122) and sp.triu(P).nnz == 375 and A.nnz == 1260, (P.shape, A.shape, sp.triu(P).nnz, A.nnz)
    np.savez_compressed(os.path.join(HERE, "scenario_mpc_small.npz"), P_data=P.data, P_indices=P.indices, P_indptr=P.indptr,
                        A_data=A.data, A_indices=A.indices, A_indptr=A.indptr, c=c, b=b, x_l=x_l, x_u=x_u, n=122, p=90)
    json.dump(GOLDEN, open(os.path.join(HERE, "scenario_mpc_small_golden.json"), "w"), indent=1)
    print("wrote fixture: n=%d p=%d nnz(P_utri)=%d nnz(A)=%d" % (P.shape[0], A.shape[0], sp.triu(P).nnz, A.nnz))
