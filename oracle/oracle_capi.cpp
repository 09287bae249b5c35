// oracle/oracle_capi.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// extern "C" surface of the CPU oracle so that tests (ctypes), smoke() and bench.py's
// cpu_baseline leg can call it.  The product (piqp_b200) never links or loads this.
#include "oracle_core.hpp"
#include "oracle_dense.hpp"
#include "oracle_sparse.hpp"
#include "oracle_multistage.hpp"
#include "oracle_sparse_cond.hpp"
#include <chrono>

using namespace oracle;

extern "C" {

// Foreign-backend hook: a table of C function pointers with the exact shape of the product's
// C-ABI (include/piqp_b200.h).  Lets the oracle's KKTSystem + IP loop act as "the reference solver"
// that calls the CUDA backend through the drop-in boundary.
struct OrcBackendVTable {
    int (*create_dense)(void** out, int n, int p, int m, const double* P_utri, const double* AT, const double* GT, int device);
    int (*create_sparse)(void** out, int n, int p, int m,
                         const int* Pp, const int* Pi, const double* Px,
                         const int* ATp, const int* ATi, const double* ATx,
                         const int* GTp, const int* GTi, const double* GTx, int mode, const int* perm, int device);
    int (*update_data)(void* h, int options, const double* P, const double* AT, const double* GT);
    int (*factor)(void* h, double delta, const double* x_reg, const double* z_reg);
    int (*solve)(void* h, const double* rx, const double* ry, const double* rz, double* lx, double* ly, double* lz);
    int (*eval_P_x)(void* h, double alpha, const double* x, double* z);
    int (*eval_A)(void* h, double an, double at, const double* xn, const double* xt, double* zn, double* zt);
    int (*eval_G)(void* h, double an, double at, const double* xn, const double* xt, double* zn, double* zt);
    void (*destroy)(void* h);
    int (*create_multistage)(void** out, int n, int p, int m,
                             const int* Pp, const int* Pi, const double* Px,
                             const int* ATp, const int* ATi, const double* ATx,
                             const int* GTp, const int* GTi, const double* GTx, int device);
};

}  // extern "C"

namespace {

struct ForeignDense : KKTBackend {
    const OrcBackendVTable* vt; void* h; const DenseMatrices& D;
    ForeignDense(const OrcBackendVTable* v, const DenseMatrices& D_) : vt(v), D(D_) {
        h = nullptr;
        vt->create_dense(&h, D.n, D.p, D.m, D.P.data(), D.AT.data(), D.GT.data(), 0);
    }
    ~ForeignDense() override { if (h) vt->destroy(h); }
    void update_data(int o) override { vt->update_data(h, o, D.P.data(), D.AT.data(), D.GT.data()); }
    bool factor(double dl, const double* xr, const double* zr) override { return vt->factor(h, dl, xr, zr) == 1; }
    void solve(const double* rx, const double* ry, const double* rz, double* lx, double* ly, double* lz) override { vt->solve(h, rx, ry, rz, lx, ly, lz); }
    void eval_P_x(double a, const double* x, double* z) override { vt->eval_P_x(h, a, x, z); }
    void eval_A(double an, double at, const double* xn, const double* xt, double* zn, double* zt) override { vt->eval_A(h, an, at, xn, xt, zn, zt); }
    void eval_G(double an, double at, const double* xn, const double* xt, double* zn, double* zt) override { vt->eval_G(h, an, at, xn, xt, zn, zt); }
};

struct ForeignSparse : KKTBackend {
    const OrcBackendVTable* vt; void* h; const SparseMatrices& S;
    ForeignSparse(const OrcBackendVTable* v, const SparseMatrices& S_) : vt(v), S(S_) {
        h = nullptr;
        if (S.kkt_solver_hint == 5 && vt->create_multistage)
            vt->create_multistage(&h, S.n, S.p, S.m, S.P.p.data(), S.P.i.data(), S.P.x.data(), S.AT.p.data(), S.AT.i.data(), S.AT.x.data(),
                                  S.GT.p.data(), S.GT.i.data(), S.GT.x.data(), 0);
        else vt->create_sparse(&h, S.n, S.p, S.m, S.P.p.data(), S.P.i.data(), S.P.x.data(), S.AT.p.data(), S.AT.i.data(), S.AT.x.data(),
                          S.GT.p.data(), S.GT.i.data(), S.GT.x.data(), (S.kkt_solver_hint >= 2 && S.kkt_solver_hint <= 4) ? S.kkt_solver_hint - 1 : 0,
                          S.user_perm.empty() ? nullptr : S.user_perm.data(), 0);
    }
    ~ForeignSparse() override { if (h) vt->destroy(h); }
    void update_data(int o) override { vt->update_data(h, o, S.P.x.data(), S.AT.x.data(), S.GT.x.data()); }
    bool factor(double dl, const double* xr, const double* zr) override { return vt->factor(h, dl, xr, zr) == 1; }
    void solve(const double* rx, const double* ry, const double* rz, double* lx, double* ly, double* lz) override { vt->solve(h, rx, ry, rz, lx, ly, lz); }
    void eval_P_x(double a, const double* x, double* z) override { vt->eval_P_x(h, a, x, z); }
    void eval_A(double an, double at, const double* xn, const double* xt, double* zn, double* zt) override { vt->eval_A(h, an, at, xn, xt, zn, zt); }
    void eval_G(double an, double at, const double* xn, const double* xt, double* zn, double* zt) override { vt->eval_G(h, an, at, xn, xt, zn, zt); }
};

std::unique_ptr<KKTBackend> make_foreign_dense(const DenseMatrices& D, void* arg) {
    return std::make_unique<ForeignDense>(static_cast<const OrcBackendVTable*>(arg), D);
}
std::unique_ptr<KKTBackend> make_foreign_sparse(const SparseMatrices& S, void* arg) {
    return std::make_unique<ForeignSparse>(static_cast<const OrcBackendVTable*>(arg), S);
}

struct Handle {
    IPSolver ip;
    OrcBackendVTable vt{};
    bool dense = true;
};

void pack(const Variables& v, const ProblemVectors& d, double* out) {
    size_t o = 0;
    auto put = [&](const Vec& a, int k) { for (int i = 0; i < k; i++) out[o++] = a[i]; };
    put(v.x, d.n); put(v.y, d.p); put(v.z_l, d.m); put(v.z_u, d.m); put(v.z_bl, d.n); put(v.z_bu, d.n);
    put(v.s_l, d.m); put(v.s_u, d.m); put(v.s_bl, d.n); put(v.s_bu, d.n);
}
void unpack(Variables& v, const ProblemVectors& d, const double* in) {
    size_t o = 0;
    v.resize(d.n, d.p, d.m);
    auto get = [&](Vec& a, int k) { for (int i = 0; i < k; i++) a[i] = in[o++]; };
    get(v.x, d.n); get(v.y, d.p); get(v.z_l, d.m); get(v.z_u, d.m); get(v.z_bl, d.n); get(v.z_bu, d.n);
    get(v.s_l, d.m); get(v.s_u, d.m); get(v.s_bl, d.n); get(v.s_bu, d.n);
}

double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

}  // namespace

extern "C" {

void orc_settings_default(Settings* s) { *s = Settings(); }
int orc_settings_size() { return (int)sizeof(Settings); }
int orc_info_size() { return (int)sizeof(Info); }

// Dense setup.  P: n x n column-major (upper triangle used); AT: n x p column-major (== A row-major);
// GT: n x m column-major (== G row-major).  Null vectors mean "not provided" (nullopt).
void* orc_dense_setup(int n, int p, int m, const double* P, const double* c, const double* AT, const double* b,
                      const double* GT, const double* h_l, const double* h_u, const double* x_l, const double* x_u,
                      const Settings* st, int identity_precond, const OrcBackendVTable* ext) {
    auto* H = new Handle();
    H->dense = true;
    if (st) H->ip.st = *st;
    auto M = std::make_unique<DenseMatrices>();
    M->resize(n, p, m);
    M->set_P(P);
    if (p > 0 && AT) M->set_AT(AT);
    if (m > 0 && GT) M->set_GT(GT);
    if (ext) { H->vt = *ext; M->backend_factory = make_foreign_dense; M->backend_factory_arg = &H->vt; }
    H->ip.pre.identity = identity_precond != 0;
    H->ip.d.resize(n, p, m);
    H->ip.M = std::move(M);
    double t0 = now_s();
    H->ip.finish_setup(c, b, h_l, h_u, x_l, x_u);
    H->ip.info.setup_time = now_s() - t0;
    return H;
}

// Dense update: null == keep.  Matrices are given UNSCALED, like the public API.
int orc_dense_update(void* h, const double* P, const double* c, const double* AT, const double* b, const double* GT,
                     const double* h_l, const double* h_u, const double* x_l, const double* x_u) {
    auto* H = static_cast<Handle*>(h);
    if (!H->ip.setup_done || !H->dense) return -1;
    auto* M = static_cast<DenseMatrices*>(H->ip.M.get());
    H->ip.begin_update();
    int opt = UPDATE_NONE;
    if (P) { M->set_P(P); opt |= UPDATE_P; }
    if (AT) { M->set_AT(AT); opt |= UPDATE_A; }
    if (GT) { M->set_GT(GT); opt |= UPDATE_G; }
    H->ip.end_update(opt, c, b, h_l, h_u, x_l, x_u);
    return 0;
}

// Sparse setup: CSC of P (upper triangle or full symmetric; only the upper part is used), of AT (n x p) and GT (n x m).
void* orc_sparse_setup(int n, int p, int m,
                       const int* Pp, const int* Pi, const double* Px,
                       const double* c,
                       const int* ATp, const int* ATi, const double* ATx, const double* b,
                       const int* GTp, const int* GTi, const double* GTx,
                       const double* h_l, const double* h_u, const double* x_l, const double* x_u,
                       const Settings* st, int identity_precond, const OrcBackendVTable* ext, const int* kkt_perm) {
    auto* H = new Handle();
    H->dense = false;
    if (st) H->ip.st = *st;
    auto M = std::make_unique<SparseMatrices>();
    M->n = n; M->p = p; M->m = m;
    M->P = Csc::upper_from(n, Pp, Pi, Px);
    M->AT = Csc::from(n, p, ATp, ATi, ATx);
    M->GT = Csc::from(n, m, GTp, GTi, GTx);
    M->kkt_solver_hint = H->ip.st.kkt_solver;
    {   // size of the KKT system of the selected mode (sparse/kkt.hpp:206-228)
        const int ks = M->kkt_solver_hint, mode = (ks >= 2 && ks <= 4) ? ks - 1 : 0;
        const int nk = n + ((mode & 1) ? 0 : p) + ((mode & 2) ? 0 : m);
        if (kkt_perm) M->user_perm.assign(kkt_perm, kkt_perm + nk);
    }
    if (ext) { H->vt = *ext; M->backend_factory = make_foreign_sparse; M->backend_factory_arg = &H->vt; }
    H->ip.pre.identity = identity_precond != 0;
    H->ip.d.resize(n, p, m);
    H->ip.M = std::move(M);
    double t0 = now_s();
    H->ip.finish_setup(c, b, h_l, h_u, x_l, x_u);
    H->ip.info.setup_time = now_s() - t0;
    return H;
}

// Sparse update with identical sparsity: value arrays only (solver.hpp:317-359). null == keep.
int orc_sparse_update(void* h, const double* Px_upper, const double* c, const double* ATx, const double* b, const double* GTx,
                      const double* h_l, const double* h_u, const double* x_l, const double* x_u) {
    auto* H = static_cast<Handle*>(h);
    if (!H->ip.setup_done || H->dense) return -1;
    auto* M = static_cast<SparseMatrices*>(H->ip.M.get());
    H->ip.begin_update();
    int opt = UPDATE_NONE;
    if (Px_upper) { std::copy(Px_upper, Px_upper + M->P.x.size(), M->P.x.begin()); opt |= UPDATE_P; }
    if (ATx) { std::copy(ATx, ATx + M->AT.x.size(), M->AT.x.begin()); opt |= UPDATE_A; }
    if (GTx) { std::copy(GTx, GTx + M->GT.x.size(), M->GT.x.begin()); opt |= UPDATE_G; }
    H->ip.end_update(opt, c, b, h_l, h_u, x_l, x_u);
    return 0;
}

int orc_solve(void* h) {
    auto* H = static_cast<Handle*>(h);
    double t0 = now_s();
    int s = H->ip.solve();
    H->ip.info.solve_time = now_s() - t0;
    return s;
}

void orc_get_info(void* h, Info* out) { *out = static_cast<Handle*>(h)->ip.info; }

// all result vectors in the public layout: x[n] y[p] z_l[m] z_u[m] z_bl[n] z_bu[n] s_l[m] s_u[m] s_bl[n] s_bu[n]
void orc_get_result(void* h, double* out) { auto* H = static_cast<Handle*>(h); pack(H->ip.it, H->ip.d, out); }

int orc_get_trace(void* h, double* out, int cap) {
    auto* H = static_cast<Handle*>(h);
    int k = std::min<int>(cap, (int)H->ip.trace.size());
    for (int i = 0; i < k; i++) out[i] = H->ip.trace[i];
    return (int)H->ip.trace.size();
}

// scaled problem data as the backend sees it (for handing identical inputs to the CUDA backend tests)
void orc_get_dims(void* h, int* out) {
    auto* H = static_cast<Handle*>(h); const auto& d = H->ip.d;
    out[0] = d.n; out[1] = d.p; out[2] = d.m; out[3] = d.n_h_l; out[4] = d.n_h_u; out[5] = d.n_x_l; out[6] = d.n_x_u;
}
void orc_dense_get_scaled(void* h, double* P, double* AT, double* GT) {
    auto* H = static_cast<Handle*>(h); auto* M = static_cast<DenseMatrices*>(H->ip.M.get());
    if (P) std::copy(M->P.begin(), M->P.end(), P);
    if (AT) std::copy(M->AT.begin(), M->AT.end(), AT);
    if (GT) std::copy(M->GT.begin(), M->GT.end(), GT);
}
// scaled sparse data as the backend sees it: values of P_utri / AT / GT in CSC order (patterns: orc_sparse_get_pattern)
void orc_sparse_get_scaled(void* h, double* Px, double* ATx, double* GTx) {
    auto* H = static_cast<Handle*>(h); auto* M = static_cast<SparseMatrices*>(H->ip.M.get());
    if (Px) std::copy(M->P.x.begin(), M->P.x.end(), Px);
    if (ATx) std::copy(M->AT.x.begin(), M->AT.x.end(), ATx);
    if (GTx) std::copy(M->GT.x.begin(), M->GT.x.end(), GTx);
}
void orc_sparse_get_nnz(void* h, int* out) {
    auto* H = static_cast<Handle*>(h); auto* M = static_cast<SparseMatrices*>(H->ip.M.get());
    out[0] = M->P.nnz(); out[1] = M->AT.nnz(); out[2] = M->GT.nnz();
}
void orc_sparse_get_pattern(void* h, int which, int* colptr, int* rowidx) {
    auto* H = static_cast<Handle*>(h); auto* M = static_cast<SparseMatrices*>(H->ip.M.get());
    const Csc& A = which == 0 ? M->P : (which == 1 ? M->AT : M->GT);
    std::copy(A.p.begin(), A.p.end(), colptr); std::copy(A.i.begin(), A.i.end(), rowidx);
}
void orc_get_scaled_vectors(void* h, double* c, double* b, double* h_l, double* h_u, double* x_l, double* x_u, double* x_b_scaling,
                            double* delta, double* delta_b, double* c_scale) {
    auto* H = static_cast<Handle*>(h); const auto& d = H->ip.d;
    auto cp = [](const Vec& a, double* o) { if (o) std::copy(a.begin(), a.end(), o); };
    cp(d.c, c); cp(d.b, b); cp(d.h_l, h_l); cp(d.h_u, h_u); cp(d.x_l, x_l); cp(d.x_u, x_u); cp(d.x_b_scaling, x_b_scaling);
    cp(H->ip.pre.delta, delta); cp(H->ip.pre.delta_b, delta_b);
    if (c_scale) *c_scale = H->ip.pre.c;
}

// KKTSystem round trip, the reference's DenseKKTTest.FactorizeSolve / SparseKKTTest.FactorizeSolve
// (tests/src/dense/kkt_test.cpp:67-139): factor at (rho, delta) with the given scaling iterate, solve rhs,
// then multiply back.  All Variables are packed as in orc_get_result but with COMPACT box blocks.
int orc_kktsystem_roundtrip(void* h, double rho, double delta, int iterative_refinement, const double* scaling,
                            const double* rhs, double* lhs, double* rhs_back) {
    auto* H = static_cast<Handle*>(h); auto& ip = H->ip;
    Variables sc, r, l, rb;
    unpack(sc, ip.d, scaling); unpack(r, ip.d, rhs);
    l.resize(ip.d.n, ip.d.p, ip.d.m); rb.resize(ip.d.n, ip.d.p, ip.d.m);
    bool ok = ip.kkt.update_scalings_and_factor(ip.d, ip.st, iterative_refinement != 0, rho, delta, sc);
    if (!ok) return 0;
    ip.kkt.solve(ip.d, ip.st, r, l);
    ip.kkt.mul(ip.d, l, rb);
    pack(l, ip.d, lhs); pack(rb, ip.d, rhs_back);
    return 1;
}

// direct access to the backend of a set-up solver (the 7 KKTSolverBase calls)
int orc_backend_factor(void* h, double delta, const double* x_reg, const double* z_reg) { return static_cast<Handle*>(h)->ip.kkt.be->factor(delta, x_reg, z_reg) ? 1 : 0; }
void orc_backend_solve(void* h, const double* rx, const double* ry, const double* rz, double* lx, double* ly, double* lz) { static_cast<Handle*>(h)->ip.kkt.be->solve(rx, ry, rz, lx, ly, lz); }
void orc_backend_eval_P_x(void* h, double a, const double* x, double* z) { static_cast<Handle*>(h)->ip.kkt.be->eval_P_x(a, x, z); }
void orc_backend_eval_A(void* h, double an, double at, const double* xn, const double* xt, double* zn, double* zt) { static_cast<Handle*>(h)->ip.kkt.be->eval_A(an, at, xn, xt, zn, zt); }
void orc_backend_eval_G(void* h, double an, double at, const double* xn, const double* xt, double* zn, double* zt) { static_cast<Handle*>(h)->ip.kkt.be->eval_G(an, at, xn, xt, zn, zt); }
// dense only: copy out the assembled lower-triangular KKT matrix and its Cholesky factor
int orc_dense_get_kkt(void* h, double* kkt, double* L) {
    auto* H = static_cast<Handle*>(h);
    auto* be = dynamic_cast<DenseKKT*>(H->ip.kkt.be.get());
    if (!be) return -1;
    if (kkt) std::copy(be->kkt.begin(), be->kkt.end(), kkt);
    if (L) std::copy(be->L.begin(), be->L.end(), L);
    return 0;
}

// multistage backend: detected block structure as (start, diag, off) triples; returns the number of blocks (incl. arrow)
int orc_multistage_blocks(void* h, int* out, int cap) {
    auto* H = static_cast<Handle*>(h);
    auto* be = dynamic_cast<MultistageKKT*>(H->ip.kkt.be.get());
    if (!be) return -1;
    int k = 0;
    for (const auto& b : be->bi) { if (3 * k + 2 < cap) { out[3 * k] = b.start; out[3 * k + 1] = b.diag; out[3 * k + 2] = b.off; } k++; }
    return k;
}
double orc_multistage_factor_flops(void* h) {
    auto* be = dynamic_cast<MultistageKKT*>(static_cast<Handle*>(h)->ip.kkt.be.get());
    return be ? be->factor_flops() : -1.0;
}
// sparse_ldlt backend: nnz(L) and the flop count of the numeric factorisation
double orc_sparse_ldlt_stats(void* h, double* nnzL) {
    auto* base = static_cast<Handle*>(h)->ip.kkt.be.get();
    const SparseLDLt* f = nullptr;
    if (auto* be = dynamic_cast<SparseKKTFull*>(base)) f = &be->ldlt;
    else if (auto* bc = dynamic_cast<SparseKKTCond*>(base)) f = &bc->ldlt;
    if (!f) return -1.0;
    if (nnzL) *nnzL = (double)f->Lp.back();
    return f->flops();
}

void orc_destroy(void* h) { delete static_cast<Handle*>(h); }

// raw dense factorisations (unit tests against scipy; reference tests/src/dense/ldlt_test.cpp:22-77)
int orc_chol(double* A, int n) { return chol_blocked(A, n, n); }
void orc_chol_solve(const double* L, int n, double* x) { chol_solve(L, n, n, x); }
int orc_ldlt(double* A, int n) { std::vector<double> t(n); return ldlt_blocked(A, n, n, t.data()); }
void orc_ldlt_solve(const double* A, int n, double* x) { ldlt_solve(A, n, n, x); }

// timed batch of dense factor+solve calls on one thread, used by bench.py's cpu_baseline leg:
// repeats (assemble + Cholesky) `reps` times and `nsolve` backend solves per factor; returns seconds.
double orc_dense_time_factor_solve(void* h, double delta, const double* x_reg, const double* z_reg,
                                   const double* rx, const double* ry, const double* rz, int reps, int nsolve,
                                   double* t_factor, double* t_solve) {
    auto* H = static_cast<Handle*>(h); auto& be = *H->ip.kkt.be; const auto& d = H->ip.d;
    Vec lx(d.n), ly(d.p), lz(d.m);
    double tf = 0, ts = 0;
    for (int r = 0; r < reps; r++) {
        double t0 = now_s();
        be.factor(delta, x_reg, z_reg);
        double t1 = now_s();
        for (int s = 0; s < nsolve; s++) be.solve(rx, ry, rz, lx.data(), ly.data(), lz.data());
        double t2 = now_s();
        tf += t1 - t0; ts += t2 - t1;
    }
    if (t_factor) *t_factor = tf;
    if (t_solve) *t_solve = ts;
    return tf + ts;
}

}  // extern "C"
