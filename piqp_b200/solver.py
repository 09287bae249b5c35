"""Batched QP solver: the reference's DenseSolver API (setup / update / solve / result / settings) with a
leading batch dimension, running the interior-point loop on the GPU (include/piqp/solver.hpp:1262-1291,
interfaces/python/src/piqp_python.cpp:130-188)."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import Info, Settings, Stats, dp

PIQP_SOLVED, PIQP_MAX_ITER_REACHED, PIQP_PRIMAL_INFEASIBLE, PIQP_DUAL_INFEASIBLE = 1, -1, -2, -3
PIQP_NUMERICS, PIQP_UNSOLVED, PIQP_INVALID_SETTINGS = -8, -9, -10


def _is_torch(a):
    return type(a).__module__.startswith("torch")


class _Arg:
    """host numpy array or CUDA torch tensor -> (pointer, on_device)"""

    def __init__(self, a, shape):
        self.keep = None
        self.ptr = None
        self.on_device = None
        if a is None:
            return
        if _is_torch(a):
            import torch
            t = a.to(torch.float64).contiguous()
            if tuple(t.shape) != tuple(shape):
                t = t.reshape(shape)
            self.keep = t
            self.ptr = C.cast(t.data_ptr(), dp)
            self.on_device = bool(t.is_cuda)
        else:
            arr = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
            if arr.shape != tuple(shape):
                arr = np.ascontiguousarray(np.broadcast_to(arr.reshape(arr.shape if arr.ndim == len(shape) else shape[1:]), shape))
            self.keep = arr
            self.ptr = arr.ctypes.data_as(dp)
            self.on_device = False


class BatchResult:
    pass


class _BatchedBase:
    def __del__(self):
        if getattr(self, "_h", None) and self._h.value:
            self._L.b200qp_cleanup(self._h)
            self._h = C.c_void_p()

    def solve(self):
        _lib.check(self._L.b200qp_update_settings(self._h, C.byref(self.settings)), "b200qp_update_settings")
        _lib.check(self._L.b200qp_solve(self._h), "b200qp_solve")
        return self.info()

    def info(self):
        arr = (Info * self.batch)()
        _lib.check(self._L.b200qp_get_info(self._h, arr), "b200qp_get_info")
        return list(arr)

    def stats(self):
        s = Stats()
        _lib.check(self._L.b200qp_get_stats(self._h, C.byref(s)), "b200qp_get_stats")
        return s

    def result(self, fields=("x", "y", "z_l", "z_u", "z_bl", "z_bu", "s_l", "s_u", "s_bl", "s_bu")):
        B, n, p, m = self.batch, self.n, self.p, self.m
        sizes = dict(x=n, y=p, z_l=m, z_u=m, z_bl=n, z_bu=n, s_l=m, s_u=m, s_bl=n, s_bu=n)
        r = BatchResult()
        ptrs = []
        for k in ("x", "y", "z_l", "z_u", "z_bl", "z_bu", "s_l", "s_u", "s_bl", "s_bu"):
            if k in fields:
                arr = np.zeros((B, sizes[k]))
                setattr(r, k, arr)
                ptrs.append(arr.ctypes.data_as(dp))
            else:
                ptrs.append(None)
        _lib.check(self._L.b200qp_get_result(self._h, *ptrs, 0), "b200qp_get_result")
        r.info = self.info()
        return r

    def work(self):
        """algorithmic (factor_flops, factor_bytes, solve_flops, solve_bytes) per backend call and instance"""
        v = [C.c_double() for _ in range(4)]
        _lib.check(self._L.b200qp_get_work(self._h, *[C.byref(x) for x in v]), "b200qp_get_work")
        return tuple(x.value for x in v)

    def set_profiling(self, enable=True):
        _lib.check(self._L.b200qp_set_profiling(self._h, int(enable)), "b200qp_set_profiling")

    def trace(self, b=0):
        rows = np.zeros((self.settings.max_iter + 1, 10))
        k = self._L.b200qp_get_trace(self._h, b, rows.ctypes.data_as(dp), rows.shape[0])
        return rows[:max(k, 0)]

    def bench_factor_solve(self, reps=3, nsolve=2):
        f = C.c_double(); s = C.c_double()
        _lib.check(self._L.b200qp_bench_factor_solve(self._h, reps, nsolve, C.byref(f), C.byref(s)), "b200qp_bench_factor_solve")
        return f.value, s.value


class SparseSolverBatched(_BatchedBase):
    """`batch` QPs that share the sparsity patterns of P, A, G (piqp::SparseSolver with a batch dimension).

    setup(P, c, A, b, G, h_l, h_u, x_l, x_u): P/A/G are scipy sparse matrices giving the shared PATTERN (their values
    are used for every instance unless Px/Ax/Gx value arrays of shape (batch, nnz) -- in CSC order -- are passed).
    """

    def __init__(self, device=0, kkt_solver="sparse_multistage"):
        self._L = _lib.lib()
        self._h = C.c_void_p()
        self.settings = Settings()
        self._L.b200qp_set_default_settings_sparse(C.byref(self.settings))
        self.settings.kkt_solver = {"sparse_ldlt": 1, "sparse_ldlt_eq_cond": 2, "sparse_ldlt_ineq_cond": 3, "sparse_ldlt_cond": 4,
                                    "sparse_multistage": 5}[kkt_solver]      # piqp::KKTSolver (settings.hpp:18-26)
        self.device = device
        self.batch = self.n = self.p = self.m = 0

    @staticmethod
    def _csc(M):
        import scipy.sparse as sp
        M = sp.csc_matrix(M)
        M.sort_indices()
        return (np.ascontiguousarray(M.indptr, dtype=np.int32), np.ascontiguousarray(M.indices, dtype=np.int32), np.ascontiguousarray(M.data, dtype=np.float64))

    def _vals(self, M_data, override, nnz, like=None):
        if override is not None:
            return _Arg(override, (self.batch, nnz))
        if like is not None:   # the other inputs live on a GPU: replicate the pattern's values there
            import torch
            return _Arg(torch.from_numpy(np.ascontiguousarray(M_data)).to(like.device).unsqueeze(0).expand(self.batch, nnz).contiguous(), (self.batch, nnz))
        return _Arg(np.broadcast_to(M_data, (self.batch, nnz)), (self.batch, nnz))

    def setup(self, batch, P, c, A=None, b=None, G=None, h_l=None, h_u=None, x_l=None, x_u=None, Px=None, Ax=None, Gx=None, kkt_perm=None):
        """kkt_perm: optional fill-reducing ordering of the KKT of the selected mode (e.g. the one rank 0 computed and broadcast)"""
        ipp = lambda a: a.ctypes.data_as(_lib.ip)
        self.batch, self.n = int(batch), P.shape[0]
        self.p = 0 if A is None else A.shape[0]
        self.m = 0 if G is None else G.shape[0]
        B, n, p, m = self.batch, self.n, self.p, self.m
        self._P = self._csc(P); self._A = self._csc(A) if p else None; self._G = self._csc(G) if m else None
        vec = lambda v, k: _Arg(None, ()) if v is None else _Arg(np.broadcast_to(np.asarray(v, dtype=np.float64), (B, k)) if not _is_torch(v) else v, (B, k))
        like = next((v for v in (Px, c, Ax, b, Gx, h_l, h_u, x_l, x_u) if v is not None and _is_torch(v) and v.is_cuda), None)
        a = [self._vals(self._P[2], Px, len(self._P[2]), like), vec(c, n), self._vals(self._A[2], Ax, len(self._A[2]), like) if p else _Arg(None, ()), vec(b, p) if p else _Arg(None, ()),
             self._vals(self._G[2], Gx, len(self._G[2]), like) if m else _Arg(None, ()), vec(h_l, m) if m else _Arg(None, ()), vec(h_u, m) if m else _Arg(None, ()), vec(x_l, n), vec(x_u, n)]
        devs = {x.on_device for x in a if x.ptr is not None}
        if len(devs) > 1:
            raise ValueError("mixing host and device inputs is not supported")
        on_dev = int(bool(devs.pop())) if devs else 0
        self._keep = a
        if self._h.value:
            self._L.b200qp_cleanup(self._h)
            self._h = C.c_void_p()
        self._perm_keep = None if kkt_perm is None else np.ascontiguousarray(kkt_perm, dtype=np.int32)
        _lib.check(self._L.b200qp_setup_sparse_ex(C.byref(self._h), B, n, p, m, ipp(self._P[0]), ipp(self._P[1]), a[0].ptr, a[1].ptr,
                                                  ipp(self._A[0]) if p else None, ipp(self._A[1]) if p else None, a[2].ptr, a[3].ptr,
                                                  ipp(self._G[0]) if m else None, ipp(self._G[1]) if m else None, a[4].ptr, a[5].ptr, a[6].ptr, a[7].ptr, a[8].ptr,
                                                  C.byref(self.settings), self.device, on_dev, None if self._perm_keep is None else ipp(self._perm_keep)), "b200qp_setup_sparse_ex")

    def kkt_perm(self):
        """the fill-reducing ordering in use (sparse_ldlt family)"""
        nk = _lib.check(self._L.b200qp_get_sparse_perm(self._h, None, 0), "b200qp_get_sparse_perm")
        out = np.zeros(nk, dtype=np.int32)
        _lib.check(self._L.b200qp_get_sparse_perm(self._h, out.ctypes.data_as(_lib.ip), nk), "b200qp_get_sparse_perm")
        return out

    def update(self, Px=None, c=None, Ax=None, b=None, Gx=None, h_l=None, h_u=None, x_l=None, x_u=None):
        B, n, p, m = self.batch, self.n, self.p, self.m
        arg = lambda v, k: _Arg(None, ()) if v is None else _Arg(np.broadcast_to(np.asarray(v, dtype=np.float64), (B, k)) if not _is_torch(v) else v, (B, k))
        a = [arg(Px, len(self._P[2])), arg(c, n), arg(Ax, len(self._A[2])) if p else _Arg(None, ()), arg(b, p), arg(Gx, len(self._G[2])) if m else _Arg(None, ()),
             arg(h_l, m), arg(h_u, m), arg(x_l, n), arg(x_u, n)]
        devs = {x.on_device for x in a if x.ptr is not None}
        on_dev = int(bool(devs.pop())) if devs else 0
        _lib.check(self._L.b200qp_update_settings(self._h, C.byref(self.settings)), "b200qp_update_settings")
        _lib.check(self._L.b200qp_update_sparse(self._h, *[x.ptr for x in a], on_dev), "b200qp_update_sparse")

    def block_info(self):
        buf = (C.c_int * 30000)()
        k = _lib.check(self._L.b200qp_multistage_blocks(self._h, buf, 30000), "b200qp_multistage_blocks")
        return [(buf[3 * i], buf[3 * i + 1], buf[3 * i + 2]) for i in range(k)]


class DenseSolverBatched(_BatchedBase):
    """`batch` independent dense QPs of identical shape:  min 1/2 x'Px + c'x  s.t. Ax=b, h_l<=Gx<=h_u, x_l<=x<=x_u."""

    def __init__(self, device=0):
        self._L = _lib.lib()
        self._h = C.c_void_p()
        self.settings = Settings()
        self._L.b200qp_set_default_settings_dense(C.byref(self.settings))
        self.device = device
        self.batch = self.n = self.p = self.m = 0

    def _args(self, P, c, A, b, G, h_l, h_u, x_l, x_u):
        B, n, p, m = self.batch, self.n, self.p, self.m
        a = [_Arg(P, (B, n, n)), _Arg(c, (B, n)), _Arg(A, (B, p, n)) if p else _Arg(None, ()), _Arg(b, (B, p)) if p else _Arg(None, ()),
             _Arg(G, (B, m, n)) if m else _Arg(None, ()), _Arg(h_l, (B, m)) if m else _Arg(None, ()), _Arg(h_u, (B, m)) if m else _Arg(None, ()),
             _Arg(x_l, (B, n)), _Arg(x_u, (B, n))]
        devs = {x.on_device for x in a if x.ptr is not None}
        if len(devs) > 1:
            raise ValueError("mixing host and device inputs is not supported")
        return a, int(bool(devs.pop())) if devs else 0

    def setup(self, P, c, A=None, b=None, G=None, h_l=None, h_u=None, x_l=None, x_u=None):
        shp = tuple(P.shape)
        if len(shp) != 3 or shp[1] != shp[2]:
            raise ValueError("P must be (batch, n, n)")
        self.batch, self.n = shp[0], shp[1]
        self.p = 0 if A is None else tuple(A.shape)[1]
        self.m = 0 if G is None else tuple(G.shape)[1]
        a, on_dev = self._args(P, c, A, b, G, h_l, h_u, x_l, x_u)
        if self._h.value:
            self._L.b200qp_cleanup(self._h)
            self._h = C.c_void_p()
        _lib.check(self._L.b200qp_setup_dense(C.byref(self._h), self.batch, self.n, self.p, self.m, *[x.ptr for x in a],
                                              C.byref(self.settings), self.device, on_dev), "b200qp_setup_dense")

    def update(self, P=None, c=None, A=None, b=None, G=None, h_l=None, h_u=None, x_l=None, x_u=None):
        a, on_dev = self._args(P, c, A, b, G, h_l, h_u, x_l, x_u)
        _lib.check(self._L.b200qp_update_settings(self._h, C.byref(self.settings)), "b200qp_update_settings")
        _lib.check(self._L.b200qp_update_dense(self._h, *[x.ptr for x in a], on_dev), "b200qp_update_dense")

    def result_device(self, x_out):
        """copy x into a CUDA torch tensor (batch, n) without leaving the device"""
        _lib.check(self._L.b200qp_get_result(self._h, C.cast(x_out.data_ptr(), dp), *([None] * 9), 1), "b200qp_get_result")
