"""Host-side mirror of the reference plugin interface `piqp::KKTSolverBase` over the C-ABI.

Same method names, argument meaning and error behaviour as include/piqp/kkt_solver_base.hpp:21-44:
`update_scalings_and_factor` returns True/False, nothing raises for a failed factorisation.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import dp, ip


def _f(a, k=None):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64).ravel())
    if k is not None and a.size != k:
        raise ValueError("expected %d elements, got %d" % (k, a.size))
    return a


def _p(a):
    return a.ctypes.data_as(dp)


KKT_UPDATE_NONE, KKT_UPDATE_P, KKT_UPDATE_A, KKT_UPDATE_G = 0, 1, 2, 4


class KKTSolverBase:
    """Owns one b200kkt_handle."""

    def __init__(self):
        self._L = _lib.lib()
        self._h = C.c_void_p()
        self.n = self.p = self.m = 0

    def __del__(self):
        if getattr(self, "_h", None) and self._h.value:
            self._L.b200kkt_destroy(self._h)
            self._h = C.c_void_p()

    def clone(self):
        o = object.__new__(type(self))
        o._L = self._L
        o.n, o.p, o.m = self.n, self.p, self.m
        o._h = C.c_void_p(self._L.b200kkt_clone(self._h))
        if not o._h.value:
            raise RuntimeError("b200kkt_clone failed: " + self._L.b200_last_error().decode())
        return o

    def update_scalings_and_factor(self, delta, x_reg, z_reg):
        x_reg, z_reg = _f(x_reg, self.n), _f(z_reg, self.m)
        r = self._L.b200kkt_factor(self._h, C.c_double(delta), _p(x_reg), _p(z_reg))
        _lib.check(r, "b200kkt_factor")
        return r == 1

    def solve(self, rhs_x, rhs_y, rhs_z):
        rx, ry, rz = _f(rhs_x, self.n), _f(rhs_y, self.p), _f(rhs_z, self.m)
        lx, ly, lz = np.zeros(self.n), np.zeros(self.p), np.zeros(self.m)
        _lib.check(self._L.b200kkt_solve(self._h, _p(rx), _p(ry), _p(rz), _p(lx), _p(ly), _p(lz)), "b200kkt_solve")
        return lx, ly, lz

    def eval_P_x(self, alpha, x):
        x = _f(x, self.n); z = np.zeros(self.n)
        _lib.check(self._L.b200kkt_eval_P_x(self._h, C.c_double(alpha), _p(x), _p(z)), "b200kkt_eval_P_x")
        return z

    def eval_A_xn_and_AT_xt(self, alpha_n, alpha_t, xn, xt):
        xn, xt = _f(xn, self.n), _f(xt, self.p)
        zn, zt = np.zeros(self.p), np.zeros(self.n)
        _lib.check(self._L.b200kkt_eval_A_xn_and_AT_xt(self._h, C.c_double(alpha_n), C.c_double(alpha_t), _p(xn), _p(xt), _p(zn), _p(zt)), "eval_A")
        return zn, zt

    def eval_G_xn_and_GT_xt(self, alpha_n, alpha_t, xn, xt):
        xn, xt = _f(xn, self.n), _f(xt, self.m)
        zn, zt = np.zeros(self.m), np.zeros(self.n)
        _lib.check(self._L.b200kkt_eval_G_xn_and_GT_xt(self._h, C.c_double(alpha_n), C.c_double(alpha_t), _p(xn), _p(xt), _p(zn), _p(zt)), "eval_G")
        return zn, zt

    def print_info(self):
        self._L.b200kkt_print_info(self._h)


class DenseKKT(KKTSolverBase):
    """Twin of piqp::dense::KKT<T> (include/piqp/dense/kkt.hpp:25-177).

    P_utri: (n, n) upper triangular; AT: (n, p); GT: (n, m) -- the members of dense::Data, any memory order.
    """

    def __init__(self, P_utri, AT=None, GT=None, device=0):
        super().__init__()
        P = np.asfortranarray(np.asarray(P_utri, dtype=np.float64))
        self.n = P.shape[0]
        AT = np.zeros((self.n, 0)) if AT is None else np.asarray(AT, dtype=np.float64)
        GT = np.zeros((self.n, 0)) if GT is None else np.asarray(GT, dtype=np.float64)
        self.p, self.m = AT.shape[1], GT.shape[1]
        ATf, GTf = np.asfortranarray(AT), np.asfortranarray(GT)
        _lib.check(self._L.b200kkt_dense_create(C.byref(self._h), self.n, self.p, self.m, _p(P),
                                                _p(ATf) if self.p else None, _p(GTf) if self.m else None, device),
                   "b200kkt_dense_create")

    def update_data(self, options, P_utri=None, AT=None, GT=None):
        P = None if P_utri is None else np.asfortranarray(np.asarray(P_utri, dtype=np.float64))
        A = None if AT is None else np.asfortranarray(np.asarray(AT, dtype=np.float64))
        G = None if GT is None else np.asfortranarray(np.asarray(GT, dtype=np.float64))
        _lib.check(self._L.b200kkt_update_data(self._h, int(options), None if P is None else _p(P),
                                               None if A is None else _p(A), None if G is None else _p(G)), "b200kkt_update_data")

    def internal_kkt_mat(self, factor=False):
        K = np.zeros((self.n, self.n), order="F")
        Lf = np.zeros((self.n, self.n), order="F")
        _lib.check(self._L.b200kkt_dense_get_kkt(self._h, _p(K), _p(Lf)), "b200kkt_dense_get_kkt")
        return Lf if factor else K


def _csc_arrays(M, upper=False):
    import scipy.sparse as sp
    M = sp.csc_matrix(M)
    if upper:
        M = sp.triu(M, format="csc")
    M.sort_indices()
    return (np.ascontiguousarray(M.indptr, dtype=np.int32), np.ascontiguousarray(M.indices, dtype=np.int32),
            np.ascontiguousarray(M.data, dtype=np.float64))


class MultistageKKT(KKTSolverBase):
    """Twin of piqp::sparse::MultistageKKT<T, I> (include/piqp/sparse/multistage_kkt.hpp:41-1816).

    P_utri: (n, n) sparse, upper triangle; AT: (n, p) sparse; GT: (n, m) sparse -- the members of sparse::Data.
    """

    def __init__(self, P_utri, AT=None, GT=None, device=0):
        import scipy.sparse as sp
        super().__init__()
        self.n = P_utri.shape[0]
        AT = sp.csc_matrix((self.n, 0)) if AT is None else AT
        GT = sp.csc_matrix((self.n, 0)) if GT is None else GT
        self.p, self.m = AT.shape[1], GT.shape[1]
        self._P = _csc_arrays(P_utri, upper=True); self._A = _csc_arrays(AT); self._G = _csc_arrays(GT)
        ipp = lambda a: a.ctypes.data_as(ip)
        _lib.check(self._L.b200kkt_multistage_create(C.byref(self._h), self.n, self.p, self.m, ipp(self._P[0]), ipp(self._P[1]), _p(self._P[2]),
                                                     ipp(self._A[0]), ipp(self._A[1]), _p(self._A[2]), ipp(self._G[0]), ipp(self._G[1]), _p(self._G[2]), device),
                   "b200kkt_multistage_create")

    def update_data(self, options, P_utri=None, AT=None, GT=None):
        P = None if P_utri is None else _csc_arrays(P_utri, upper=True)[2]
        A = None if AT is None else _csc_arrays(AT)[2]
        G = None if GT is None else _csc_arrays(GT)[2]
        _lib.check(self._L.b200kkt_update_data(self._h, int(options), None if P is None else _p(P), None if A is None else _p(A),
                                               None if G is None else _p(G)), "b200kkt_update_data")

    def block_info(self):
        buf = (C.c_int * 30000)()
        k = _lib.check(self._L.b200kkt_multistage_blocks(self._h, buf, 30000), "b200kkt_multistage_blocks")
        return [(buf[3 * i], buf[3 * i + 1], buf[3 * i + 2]) for i in range(k)]


class SparseKKT(KKTSolverBase):
    """Twin of piqp::sparse::KKT<T, I, Mode> (include/piqp/sparse/kkt.hpp:31-250): LDL^T of the permuted
    quasi-definite KKT matrix.  mode = KKTMode (kkt_fwd.hpp:15-21): 0 FULL, 1 EQ_ELIMINATED, 2 INEQ_ELIMINATED,
    3 ALL_ELIMINATED.  `perm` (optional, length of that mode's KKT, perm[new] = old) replaces the built-in ordering."""

    MODES = {"sparse_ldlt": 0, "sparse_ldlt_eq_cond": 1, "sparse_ldlt_ineq_cond": 2, "sparse_ldlt_cond": 3}

    def __init__(self, P_utri, AT=None, GT=None, perm=None, device=0, mode=0):
        import scipy.sparse as sp
        super().__init__()
        self.n = P_utri.shape[0]
        AT = sp.csc_matrix((self.n, 0)) if AT is None else AT
        GT = sp.csc_matrix((self.n, 0)) if GT is None else GT
        self.p, self.m = AT.shape[1], GT.shape[1]
        self._P = _csc_arrays(P_utri, upper=True); self._A = _csc_arrays(AT); self._G = _csc_arrays(GT)
        ipp = lambda a: a.ctypes.data_as(ip)
        pm = None if perm is None else np.ascontiguousarray(perm, dtype=np.int32)
        self.mode = self.MODES[mode] if isinstance(mode, str) else int(mode)
        self.n_kkt = self.n + (0 if self.mode & 1 else self.p) + (0 if self.mode & 2 else self.m)
        _lib.check(self._L.b200kkt_sparse_create(C.byref(self._h), self.n, self.p, self.m, ipp(self._P[0]), ipp(self._P[1]), _p(self._P[2]),
                                                 ipp(self._A[0]), ipp(self._A[1]), _p(self._A[2]), ipp(self._G[0]), ipp(self._G[1]), _p(self._G[2]),
                                                 self.mode, None if pm is None else ipp(pm), device), "b200kkt_sparse_create")

    update_data = None  # set below (shared with MultistageKKT)

    def symbolic_info(self):
        nk = self.n_kkt
        a, b, lv = C.c_longlong(), C.c_longlong(), C.c_int()
        perm = np.zeros(nk, dtype=np.int32)
        _lib.check(self._L.b200kkt_sparse_info(self._h, C.byref(a), C.byref(b), C.byref(lv), perm.ctypes.data_as(ip)), "b200kkt_sparse_info")
        return {"nnz_kkt": a.value, "nnz_L": b.value, "etree_levels": lv.value, "perm": perm}


SparseKKT.update_data = MultistageKKT.update_data


class LDLTNoPivot:
    """Twin of piqp::dense::LDLTNoPivot<Mat, UpLo> (include/piqp/dense/ldlt_no_pivot.hpp:276-355 compute, :393-450 solveInPlace;
    SURVEY a11): LDL^T without pivoting of a dense symmetric (quasi-)definite matrix.  The dense matrix is one supernode of
    the sparse multifrontal backend -- a single front of n rows -- so it runs on the same kernels: the 64-column panel /
    DMMA trailing-update blocked LDL^T and the blocked triangular solves of the whole-GPU schedule for n >= 512
    (piqp_b200/csrc/sparse_wide.cuh), the shared-memory / HBM front kernels below that.  Not on the solver's call path."""

    def __init__(self, device=0):
        self.device, self._kkt, self._ok, self.n = device, None, False, 0

    def compute(self, P, uplo="lower"):
        """P: (n, n) array whose `uplo` triangle holds the symmetric matrix (the other triangle is ignored)"""
        import scipy.sparse as sp
        P = np.asarray(P, dtype=np.float64)
        n = P.shape[0]
        U = np.triu(P) if uplo == "upper" else np.tril(P).T
        # the FULL upper pattern, zeros stored explicitly: one supernode of width n whatever the values are
        indptr = np.cumsum([0] + [j + 1 for j in range(n)]).astype(np.int32)
        indices = np.concatenate([np.arange(j + 1) for j in range(n)]).astype(np.int32)
        data = np.concatenate([U[:j + 1, j] for j in range(n)])
        Pu = sp.csc_matrix((data, indices, indptr), shape=(n, n))
        if self._kkt is None or self.n != n:
            self._kkt = SparseKKT(Pu, perm=np.arange(n, dtype=np.int32), device=self.device, mode=0)
            self.n = n
        else:
            self._kkt.update_data(1, Pu, None, None)
        self._ok = bool(self._kkt.update_scalings_and_factor(1.0, np.zeros(n), np.zeros(0)))
        return self

    def info(self):
        """True = Eigen::Success"""
        return self._ok

    def solve(self, b):
        x, _, _ = self._kkt.solve(np.asarray(b, dtype=np.float64), np.zeros(0), np.zeros(0))
        return x


def sparse_ldlt_symbolic(P, A, G, perm=None, mode=0):
    """Host-only symbolic phase of the sparse_ldlt backend (b200_sparse_ldlt_symbolic_mode): fill-reducing ordering of the
    KKT of the given KKTMode (sparse/ordering.hpp:59-125), nnz(L), etree levels and factor flops (sparse/ldlt.hpp:42-99),
    supernode count and largest front of the multifrontal schedule.
    P (n x n, upper part used), A (p x n) or None, G (m x n) or None: scipy sparse."""
    import scipy.sparse as sp
    L = _lib.lib()
    n = P.shape[0]
    AT = sp.csc_matrix((n, 0)) if A is None else sp.csc_matrix(sp.csc_matrix(A).T)
    GT = sp.csc_matrix((n, 0)) if G is None else sp.csc_matrix(sp.csc_matrix(G).T)
    Pp, Pi, _ = _csc_arrays(P, upper=True)
    Ap, Ai, _ = _csc_arrays(AT)
    Gp, Gi, _ = _csc_arrays(GT)
    p, m = AT.shape[1], GT.shape[1]
    mode = SparseKKT.MODES[mode] if isinstance(mode, str) else int(mode)
    out = np.zeros(n + (0 if mode & 1 else p) + (0 if mode & 2 else m), dtype=np.int32)
    nk, nl, lv, fl, ns, fm = C.c_longlong(), C.c_longlong(), C.c_int(), C.c_double(), C.c_int(), C.c_int()
    pin = None if perm is None else np.ascontiguousarray(perm, dtype=np.int32)
    q = lambda a: a.ctypes.data_as(ip)
    _lib.check(L.b200_sparse_ldlt_symbolic_mode(n, p, m, q(Pp), q(Pi), q(Ap), q(Ai), q(Gp), q(Gi), mode, None if pin is None else q(pin), q(out),
                                                C.byref(nk), C.byref(nl), C.byref(lv), C.byref(fl), C.byref(ns), C.byref(fm)), "b200_sparse_ldlt_symbolic_mode")
    return {"perm": out, "nnz_kkt": nk.value, "nnz_L": nl.value, "levels": lv.value, "flops": fl.value, "supernodes": ns.value, "largest_front": fm.value}


def c_abi_vtable():
    """Function-pointer table of the C-ABI in the layout oracle/oracle_capi.cpp::OrcBackendVTable expects
    (used by tests to put the CUDA backend behind the oracle's KKTSystem + IP loop)."""
    L = _lib.lib()

    def addr(name):
        return C.cast(getattr(L, name), C.c_void_p).value

    return {
        "create_dense": addr("b200kkt_dense_create"), "create_sparse": addr("b200kkt_sparse_create"),
        "update_data": addr("b200kkt_update_data"), "factor": addr("b200kkt_factor"), "solve": addr("b200kkt_solve"),
        "eval_P_x": addr("b200kkt_eval_P_x"), "eval_A": addr("b200kkt_eval_A_xn_and_AT_xt"),
        "eval_G": addr("b200kkt_eval_G_xn_and_GT_xt"), "destroy": addr("b200kkt_destroy"),
        "create_multistage": addr("b200kkt_multistage_create"),
    }
