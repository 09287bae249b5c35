"""ctypes binding of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE, NOT PRODUCT CODE: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.  The product package (piqp_b200)
never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)


class Settings(C.Structure):
    """Mirror of oracle::Settings == piqp::Settings (include/piqp/settings.hpp:43-82)."""
    _fields_ = [
        ("rho_init", C.c_double), ("delta_init", C.c_double),
        ("eps_abs", C.c_double), ("eps_rel", C.c_double),
        ("check_duality_gap", C.c_int),
        ("eps_duality_gap_abs", C.c_double), ("eps_duality_gap_rel", C.c_double),
        ("infeasibility_threshold", C.c_double),
        ("reg_lower_limit", C.c_double), ("reg_finetune_lower_limit", C.c_double),
        ("reg_finetune_primal_update_threshold", C.c_long), ("reg_finetune_dual_update_threshold", C.c_long),
        ("max_iter", C.c_long), ("max_factor_retires", C.c_long),
        ("preconditioner_scale_cost", C.c_int), ("preconditioner_reuse_on_update", C.c_int),
        ("preconditioner_iter", C.c_long),
        ("tau", C.c_double),
        ("kkt_solver", C.c_int),
        ("iterative_refinement_always_enabled", C.c_int),
        ("iterative_refinement_eps_abs", C.c_double), ("iterative_refinement_eps_rel", C.c_double),
        ("iterative_refinement_max_iter", C.c_long),
        ("iterative_refinement_min_improvement_rate", C.c_double),
        ("iterative_refinement_static_regularization_eps", C.c_double),
        ("iterative_refinement_static_regularization_rel", C.c_double),
        ("verbose", C.c_int), ("compute_timings", C.c_int),
    ]


class Info(C.Structure):
    """Mirror of oracle::Info == piqp::Info (include/piqp/results.hpp:45-89) + 3 counters."""
    _fields_ = [
        ("status", C.c_int), ("iter", C.c_long),
        ("rho", C.c_double), ("delta", C.c_double), ("mu", C.c_double), ("sigma", C.c_double),
        ("primal_step", C.c_double), ("dual_step", C.c_double),
        ("primal_res", C.c_double), ("primal_res_rel", C.c_double), ("dual_res", C.c_double), ("dual_res_rel", C.c_double),
        ("primal_res_reg", C.c_double), ("primal_res_reg_rel", C.c_double),
        ("dual_res_reg", C.c_double), ("dual_res_reg_rel", C.c_double),
        ("primal_prox_inf", C.c_double), ("dual_prox_inf", C.c_double),
        ("prev_primal_res", C.c_double), ("prev_dual_res", C.c_double),
        ("primal_obj", C.c_double), ("dual_obj", C.c_double), ("duality_gap", C.c_double), ("duality_gap_rel", C.c_double),
        ("factor_retires", C.c_long), ("reg_limit", C.c_double),
        ("no_primal_update", C.c_long), ("no_dual_update", C.c_long),
        ("setup_time", C.c_double), ("update_time", C.c_double), ("solve_time", C.c_double),
        ("kkt_factor_time", C.c_double), ("kkt_solve_time", C.c_double), ("run_time", C.c_double),
        ("n_factor", C.c_long), ("n_solve", C.c_long), ("n_backend_solve", C.c_long),
    ]


class BackendVTable(C.Structure):
    """Function-pointer table with the shape of the product's C-ABI (include/piqp_b200.h)."""
    _fields_ = [(name, C.c_void_p) for name in
                ("create_dense", "create_sparse", "update_data", "factor", "solve",
                 "eval_P_x", "eval_A", "eval_G", "destroy", "create_multistage")]


STATUS = {1: "solved", -1: "max_iter_reached", -2: "primal_infeasible", -3: "dual_infeasible",
          -8: "numerics", -9: "unsolved", -10: "invalid_settings"}
KKT_SOLVERS = {"dense_cholesky": 0, "sparse_ldlt": 1, "sparse_ldlt_eq_cond": 2, "sparse_ldlt_ineq_cond": 3,
               "sparse_ldlt_cond": 4, "sparse_multistage": 5}


def build(native=False, force=False):
    """Compile the oracle with g++ (seconds).  native=True builds liboracle_native.so for timing."""
    out = "liboracle_native.so" if native else "liboracle.so"
    path = os.path.join(_HERE, out)
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".hpp", ".cpp")) or f == "Makefile"]
    if force or not os.path.exists(path) or any(os.path.getmtime(s) > os.path.getmtime(path) for s in srcs):
        args = ["make", "-C", _HERE, "OUT=" + out] + (["ARCH=native"] if native else [])
        subprocess.run(args, check=True, capture_output=True)
    return path


def lib(native=False):
    global _LIB
    key = "native" if native else "portable"
    if _LIB is None:
        _LIB = {}
    if key in _LIB:
        return _LIB[key]
    path = os.path.join(_HERE, "liboracle_native.so" if native else "liboracle.so")
    if not os.path.exists(path):
        try:
            build(native=native)
        except Exception as e:  # no compiler on this box: fall back to the shipped portable build
            if native:
                return lib(False)
            raise RuntimeError("oracle library missing and cannot be built: %s" % e)
    L = C.CDLL(path)
    L.orc_dense_setup.restype = C.c_void_p
    L.orc_sparse_setup.restype = C.c_void_p
    L.orc_dense_time_factor_solve.restype = C.c_double
    L.orc_multistage_factor_flops.restype = C.c_double
    L.orc_sparse_ldlt_stats.restype = C.c_double
    L.orc_multistage_blocks.restype = C.c_int
    for f in ("orc_solve", "orc_dense_update", "orc_sparse_update", "orc_get_trace", "orc_kktsystem_roundtrip",
              "orc_backend_factor", "orc_dense_get_kkt", "orc_chol", "orc_ldlt", "orc_settings_size", "orc_info_size"):
        getattr(L, f).restype = C.c_int
    assert L.orc_settings_size() == C.sizeof(Settings), (L.orc_settings_size(), C.sizeof(Settings))
    assert L.orc_info_size() == C.sizeof(Info), (L.orc_info_size(), C.sizeof(Info))
    _LIB[key] = L
    return L


def _dp(a):
    return None if a is None else a.ctypes.data_as(c_double_p)


def _ip(a):
    return None if a is None else a.ctypes.data_as(c_int_p)


def _vec(a, k=None):
    if a is None:
        return None
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64).ravel())
    a = np.where(np.isposinf(a), 1e30, a)
    a = np.where(np.isneginf(a), -1e30, a)
    a = np.ascontiguousarray(a)
    if k is not None:
        assert a.size == k, (a.size, k)
    return a


def default_settings(**kw):
    s = Settings()
    lib().orc_settings_default(C.byref(s))
    for k, v in kw.items():
        if k == "kkt_solver" and isinstance(v, str):
            v = KKT_SOLVERS[v]
        if not hasattr(s, k):
            raise AttributeError(k)
        setattr(s, k, v)
    return s


class Result:
    pass


class _Base:
    def __init__(self, native=False):
        self._L = lib(native)
        self._h = None
        self._keep = []

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.orc_destroy(C.c_void_p(self._h))
            self._h = None

    @property
    def dims(self):
        out = (C.c_int * 7)()
        self._L.orc_get_dims(C.c_void_p(self._h), out)
        return tuple(out)

    def solve(self):
        st = self._L.orc_solve(C.c_void_p(self._h))
        return st

    def info(self):
        i = Info()
        self._L.orc_get_info(C.c_void_p(self._h), C.byref(i))
        return i

    def result(self):
        n, p, m = self.dims[:3]
        buf = np.zeros(5 * n + p + 4 * m)
        self._L.orc_get_result(C.c_void_p(self._h), _dp(buf))
        r = Result()
        o = 0
        for name, k in (("x", n), ("y", p), ("z_l", m), ("z_u", m), ("z_bl", n), ("z_bu", n),
                        ("s_l", m), ("s_u", m), ("s_bl", n), ("s_bu", n)):
            setattr(r, name, buf[o:o + k].copy())
            o += k
        r.info = self.info()
        return r

    def trace(self):
        """per-iteration rows: rho, delta, mu, primal_step, dual_step, primal_res, dual_res, primal_obj, dual_obj, gap"""
        cap = 10 * 300
        buf = np.zeros(cap)
        k = self._L.orc_get_trace(C.c_void_p(self._h), _dp(buf), cap)
        return buf[:k].reshape(-1, 10)

    def scaled_vectors(self):
        n, p, m = self.dims[:3]
        out = {k: np.zeros(sz) for k, sz in (("c", n), ("b", p), ("h_l", m), ("h_u", m), ("x_l", n), ("x_u", n),
                                             ("x_b_scaling", n), ("delta", n + p + m), ("delta_b", n))}
        cs = C.c_double()
        self._L.orc_get_scaled_vectors(C.c_void_p(self._h), *[_dp(out[k]) for k in
                                       ("c", "b", "h_l", "h_u", "x_l", "x_u", "x_b_scaling", "delta", "delta_b")], C.byref(cs))
        out["c_scale"] = cs.value
        return out

    # --- the 7 backend calls on the (scaled) data of this solver
    def backend_factor(self, delta, x_reg, z_reg):
        x_reg = _vec(x_reg); z_reg = _vec(z_reg)
        return self._L.orc_backend_factor(C.c_void_p(self._h), C.c_double(delta), _dp(x_reg), _dp(z_reg))

    def backend_solve(self, rx, ry, rz):
        n, p, m = self.dims[:3]
        rx, ry, rz = _vec(rx, n), _vec(ry, p), _vec(rz, m)
        lx, ly, lz = np.zeros(n), np.zeros(p), np.zeros(m)
        self._L.orc_backend_solve(C.c_void_p(self._h), _dp(rx), _dp(ry), _dp(rz), _dp(lx), _dp(ly), _dp(lz))
        return lx, ly, lz

    def backend_eval_P_x(self, alpha, x):
        n = self.dims[0]
        x = _vec(x, n); z = np.zeros(n)
        self._L.orc_backend_eval_P_x(C.c_void_p(self._h), C.c_double(alpha), _dp(x), _dp(z))
        return z

    def _eval_nt(self, fn, k, an, at, xn, xt):
        n = self.dims[0]
        xn = _vec(xn, n); xt = _vec(xt, k)
        zn = np.zeros(k); zt = np.zeros(n)
        fn(C.c_void_p(self._h), C.c_double(an), C.c_double(at), _dp(xn), _dp(xt), _dp(zn), _dp(zt))
        return zn, zt

    def backend_eval_A(self, an, at, xn, xt):
        return self._eval_nt(self._L.orc_backend_eval_A, self.dims[1], an, at, xn, xt)

    def backend_eval_G(self, an, at, xn, xt):
        return self._eval_nt(self._L.orc_backend_eval_G, self.dims[2], an, at, xn, xt)

    def kktsystem_roundtrip(self, rho, delta, scaling, rhs, iterative_refinement=False):
        n, p, m = self.dims[:3]
        N = 5 * n + p + 4 * m
        scaling = _vec(scaling, N); rhs = _vec(rhs, N)
        lhs = np.zeros(N); back = np.zeros(N)
        ok = self._L.orc_kktsystem_roundtrip(C.c_void_p(self._h), C.c_double(rho), C.c_double(delta),
                                             int(iterative_refinement), _dp(scaling), _dp(rhs), _dp(lhs), _dp(back))
        return ok, lhs, back


class DenseSolver(_Base):
    """Oracle twin of piqp::DenseSolver (include/piqp/solver.hpp:1262-1291)."""

    def __init__(self, settings=None, identity_preconditioner=False, backend_vtable=None, native=False):
        super().__init__(native)
        self.settings = settings if settings is not None else default_settings()
        self.identity = identity_preconditioner
        self.vt = backend_vtable

    def setup(self, P, c, A=None, b=None, G=None, h_l=None, h_u=None, x_l=None, x_u=None):
        P = np.asarray(P, dtype=np.float64)
        n = P.shape[0]
        p = 0 if A is None else np.asarray(A).shape[0]
        m = 0 if G is None else np.asarray(G).shape[0]
        Pf = np.asfortranarray(P)
        AT = None if p == 0 else np.ascontiguousarray(np.asarray(A, dtype=np.float64))   # A row-major == AT col-major
        GT = None if m == 0 else np.ascontiguousarray(np.asarray(G, dtype=np.float64))
        args = [Pf, _vec(c, n), AT, _vec(b, p) if p else None, GT, _vec(h_l, m) if m and h_l is not None else None,
                _vec(h_u, m) if m and h_u is not None else None, _vec(x_l, n), _vec(x_u, n)]
        self._keep = args
        if self._h:
            self._L.orc_destroy(C.c_void_p(self._h))
        self._h = self._L.orc_dense_setup(n, p, m, _dp(Pf), _dp(args[1]), _dp(AT), _dp(args[3]), _dp(GT), _dp(args[5]),
                                          _dp(args[6]), _dp(args[7]), _dp(args[8]), C.byref(self.settings),
                                          int(self.identity), C.byref(self.vt) if self.vt is not None else None)

    def update(self, P=None, c=None, A=None, b=None, G=None, h_l=None, h_u=None, x_l=None, x_u=None):
        n, p, m = self.dims[:3]
        Pf = None if P is None else np.asfortranarray(np.asarray(P, dtype=np.float64))
        AT = None if A is None else np.ascontiguousarray(np.asarray(A, dtype=np.float64))
        GT = None if G is None else np.ascontiguousarray(np.asarray(G, dtype=np.float64))
        v = [_vec(c), _vec(b), _vec(h_l), _vec(h_u), _vec(x_l), _vec(x_u)]
        return self._L.orc_dense_update(C.c_void_p(self._h), _dp(Pf), _dp(v[0]), _dp(AT), _dp(v[1]), _dp(GT),
                                        _dp(v[2]), _dp(v[3]), _dp(v[4]), _dp(v[5]))

    def scaled_matrices(self):
        n, p, m = self.dims[:3]
        P = np.zeros((n, n), order="F"); AT = np.zeros((n, p), order="F"); GT = np.zeros((n, m), order="F")
        self._L.orc_dense_get_scaled(C.c_void_p(self._h), _dp(P), _dp(AT), _dp(GT))
        return P, AT, GT

    def kkt_and_factor(self):
        n = self.dims[0]
        K = np.zeros((n, n), order="F"); Lf = np.zeros((n, n), order="F")
        self._L.orc_dense_get_kkt(C.c_void_p(self._h), _dp(K), _dp(Lf))
        return K, Lf

    def time_factor_solve(self, delta, x_reg, z_reg, rx, ry, rz, reps, nsolve):
        tf = C.c_double(); ts = C.c_double()
        a = [_vec(x) for x in (x_reg, z_reg, rx, ry, rz)]
        self._L.orc_dense_time_factor_solve(C.c_void_p(self._h), C.c_double(delta), *[_dp(x) for x in a], reps, nsolve,
                                            C.byref(tf), C.byref(ts))
        return tf.value, ts.value


def _csc(M, shape=None, upper=False):
    import scipy.sparse as sp
    M = sp.csc_matrix(M) if shape is None else sp.csc_matrix(M, shape=shape)
    if upper:
        M = sp.triu(M, format="csc")
    M.sort_indices()
    return (np.ascontiguousarray(M.indptr, dtype=np.int32), np.ascontiguousarray(M.indices, dtype=np.int32),
            np.ascontiguousarray(M.data, dtype=np.float64))


class SparseSolver(_Base):
    """Oracle twin of piqp::SparseSolver (include/piqp/solver.hpp:1293-1322)."""

    def __init__(self, settings=None, identity_preconditioner=False, backend_vtable=None, native=False, kkt_perm=None):
        super().__init__(native)
        self.settings = settings if settings is not None else default_settings(kkt_solver="sparse_ldlt")
        self.identity = identity_preconditioner
        self.vt = backend_vtable
        self.kkt_perm = kkt_perm

    def setup(self, P, c, A=None, b=None, G=None, h_l=None, h_u=None, x_l=None, x_u=None):
        import scipy.sparse as sp
        n = P.shape[0]
        p = 0 if A is None else A.shape[0]
        m = 0 if G is None else G.shape[0]
        Pp, Pi, Px = _csc(P, upper=True)
        ATp, ATi, ATx = _csc(sp.csc_matrix((n, 0)) if p == 0 else sp.csc_matrix(A).T)
        GTp, GTi, GTx = _csc(sp.csc_matrix((n, 0)) if m == 0 else sp.csc_matrix(G).T)
        v = [_vec(c, n), _vec(b, p) if p else None, _vec(h_l, m) if m and h_l is not None else None,
             _vec(h_u, m) if m and h_u is not None else None, _vec(x_l, n), _vec(x_u, n)]
        perm = None if self.kkt_perm is None else np.ascontiguousarray(self.kkt_perm, dtype=np.int32)
        self._keep = [Pp, Pi, Px, ATp, ATi, ATx, GTp, GTi, GTx, v, perm]
        self._nnz = (len(Px), len(ATx), len(GTx))
        if self._h:
            self._L.orc_destroy(C.c_void_p(self._h))
        self._h = self._L.orc_sparse_setup(n, p, m, _ip(Pp), _ip(Pi), _dp(Px), _dp(v[0]), _ip(ATp), _ip(ATi), _dp(ATx), _dp(v[1]),
                                           _ip(GTp), _ip(GTi), _dp(GTx), _dp(v[2]), _dp(v[3]), _dp(v[4]), _dp(v[5]),
                                           C.byref(self.settings), int(self.identity),
                                           C.byref(self.vt) if self.vt is not None else None, _ip(perm))

    def scaled_matrices(self):
        """(P_utri, AT, GT) as scipy CSC matrices with the Ruiz-scaled values the backend works on"""
        import scipy.sparse as sp
        n, p, m = self.dims[:3]
        nnz = (C.c_int * 3)()
        self._L.orc_sparse_get_nnz(C.c_void_p(self._h), nnz)
        vals = [np.zeros(nnz[k]) for k in range(3)]
        self._L.orc_sparse_get_scaled(C.c_void_p(self._h), _dp(vals[0]), _dp(vals[1]), _dp(vals[2]))
        out = []
        for k, (r, c_) in enumerate(((n, n), (n, p), (n, m))):
            cp = np.zeros(c_ + 1, dtype=np.int32); ri = np.zeros(max(nnz[k], 1), dtype=np.int32)
            self._L.orc_sparse_get_pattern(C.c_void_p(self._h), k, _ip(cp), _ip(ri))
            out.append(sp.csc_matrix((vals[k], ri[:nnz[k]], cp), shape=(r, c_)))
        return out

    def multistage_blocks(self):
        buf = (C.c_int * 30000)()
        k = self._L.orc_multistage_blocks(C.c_void_p(self._h), buf, 30000)
        return [(buf[3 * i], buf[3 * i + 1], buf[3 * i + 2]) for i in range(max(k, 0))]

    def ldlt_stats(self):
        nn = C.c_double()
        fl = self._L.orc_sparse_ldlt_stats(C.c_void_p(self._h), C.byref(nn))
        return nn.value, fl

    def update(self, P=None, c=None, A=None, b=None, G=None, h_l=None, h_u=None, x_l=None, x_u=None):
        import scipy.sparse as sp
        Px = None if P is None else _csc(P, upper=True)[2]
        ATx = None if A is None else _csc(sp.csc_matrix(A).T)[2]
        GTx = None if G is None else _csc(sp.csc_matrix(G).T)[2]
        v = [_vec(c), _vec(b), _vec(h_l), _vec(h_u), _vec(x_l), _vec(x_u)]
        return self._L.orc_sparse_update(C.c_void_p(self._h), _dp(Px), _dp(v[0]), _dp(ATx), _dp(v[1]), _dp(GTx),
                                         _dp(v[2]), _dp(v[3]), _dp(v[4]), _dp(v[5]))


def chol(A):
    """blocked Cholesky of the lower triangle of A (returns L, info); info = -1 on success."""
    A = np.asfortranarray(np.array(A, dtype=np.float64))
    info = lib().orc_chol(_dp(A), A.shape[0])
    return np.tril(A), info


def ldlt(A):
    """LDLTNoPivot (lower): returns (unit L, D, info)."""
    A = np.asfortranarray(np.array(A, dtype=np.float64))
    info = lib().orc_ldlt(_dp(A), A.shape[0])
    return np.tril(A, -1) + np.eye(A.shape[0]), np.diag(A).copy(), info, A


def ldlt_solve(Afac, b):
    x = _vec(b).copy()
    lib().orc_ldlt_solve(_dp(np.asfortranarray(Afac)), Afac.shape[0], _dp(x))
    return x
