// oracle/oracle_sparse.hpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement of the general sparse KKT backend of PIQP v0.6.2 (mode KKT_FULL, the default
// `sparse_ldlt`):
//   sparse::KKT<FULL>                       include/piqp/sparse/kkt.hpp:31-250
//   KKTImpl<FULL> (pattern + value maps)    include/piqp/sparse/kkt_full.hpp:39-251
//   permute_sparse_symmetric_matrix         include/piqp/sparse/utils.hpp:31-128
//   sparse::LDLt (SuiteSparse-LDL restated) include/piqp/sparse/ldlt.hpp:42-218
//   AMDOrdering perm/permt                  include/piqp/sparse/ordering.hpp:59-125
//   sparse Ruiz sweeps                      include/piqp/sparse/preconditioner.hpp:81-180
// Eigen::AMDOrdering is third-party (Eigen 3.4.0, not under /root/reference).  Any fill-reducing
// permutation yields the same solve() results to rounding, so the oracle takes the permutation from
// the caller when given (tests hand it the product's AMD so both arms factor the same matrix) and
// otherwise uses its own exact minimum-degree ordering (quotient-graph free, fine for test sizes).
#pragma once
#include "oracle_core.hpp"
#include "experimental_multifrontal.hpp"
#include <numeric>
#include <set>

namespace oracle {

struct Csc {
    int rows = 0, cols = 0;
    IVec p, i;
    Vec x;
    int nnz() const { return (int)i.size(); }
    static Csc from(int rows, int cols, const int* cp, const int* ri, const double* v) {
        Csc A; A.rows = rows; A.cols = cols; A.p.assign(cols + 1, 0);
        if (cp) { A.p.assign(cp, cp + cols + 1); A.i.assign(ri, ri + cp[cols]); A.x.assign(v, v + cp[cols]); }
        return A;
    }
    // keep only entries with row <= col (solver.hpp:182 triangularView<Upper> on a sparse matrix)
    static Csc upper_from(int n, const int* cp, const int* ri, const double* v) {
        Csc A; A.rows = n; A.cols = n; A.p.assign(n + 1, 0);
        for (int j = 0; j < n; j++) {
            for (int k = cp[j]; k < cp[j + 1]; k++) if (ri[k] <= j) { A.i.push_back(ri[k]); A.x.push_back(v[k]); }
            A.p[j + 1] = (int)A.i.size();
        }
        return A;
    }
};

// ldlt.hpp:22-218
struct SparseLDLt {
    IVec etree, Lp, Lnz, Li, flag, pattern;
    Vec Lx, D, Dinv, y;

    void symbolic(const Csc& A) {  // :42-99
        const int n = A.rows;
        etree.assign(n, -1); Lp.assign(n + 1, 0); Lnz.assign(n, 0); D.assign(n, 0); Dinv.assign(n, 0);
        flag.assign(n, 0); pattern.assign(n, 0); y.assign(n, 0);
        for (int k = 0; k < n; k++) {
            etree[k] = -1; flag[k] = k; Lnz[k] = 0;
            for (int q = A.p[k]; q < A.p[k + 1]; q++) {
                for (int i = A.i[q]; flag[i] != k; i = etree[i]) {
                    if (etree[i] == -1) etree[i] = k;
                    Lnz[i]++; flag[i] = k;
                }
            }
        }
        for (int k = 0; k < n; k++) Lp[k + 1] = Lp[k] + Lnz[k];
        Li.assign(Lp[n], 0); Lx.assign(Lp[n], 0);
    }

    // :101-169 ; returns n on success, failing row otherwise.  No FMA contraction (see Makefile: -ffp-contract=off).
    int numeric(const Csc& A) {
        const int n = A.rows;
        if (const char* e = getenv("ORACLE_MULTIFRONTAL")) if (e[0] != '0') return numeric_multifrontal(*this, A, atoi(e) - 1);   // experimental_multifrontal.hpp (study only)
        for (int k = 0; k < n; k++) {
            y[k] = 0.0; int top = n; flag[k] = k; Lnz[k] = 0;
            for (int q = A.p[k]; q < A.p[k + 1]; q++) {
                int i = A.i[q]; y[i] = A.x[q];
                int len;
                for (len = 0; flag[i] != k; i = etree[i]) { pattern[len++] = i; flag[i] = k; }
                while (len > 0) pattern[--top] = pattern[--len];
            }
            D[k] = y[k]; y[k] = 0.0;
            for (; top < n; top++) {
                const int i = pattern[top];
                const double yi = y[i]; y[i] = 0.0;
                const int q2 = Lp[i] + Lnz[i];
                int q;
                for (q = Lp[i]; q < q2; q++) { volatile double t = Lx[q] * yi; y[Li[q]] -= t; }
                const double lki = yi / D[i];
                volatile double t = lki * yi;
                D[k] -= t;
                Li[q] = k; Lx[q] = lki; Lnz[i]++;
            }
            if (D[k] == 0.0) return k;
        }
        for (int k = 0; k < n; k++) Dinv[k] = 1.0 / D[k];
        return n;
    }

    void solve_inplace(double* x) const {  // :171-218
        const int n = (int)D.size();
        for (int j = 0; j < n; j++) for (int q = Lp[j]; q < Lp[j + 1]; q++) x[Li[q]] -= Lx[q] * x[j];
        for (int j = 0; j < n; j++) x[j] *= Dinv[j];
        for (int j = n - 1; j >= 0; j--) for (int q = Lp[j]; q < Lp[j + 1]; q++) x[j] -= Lx[q] * x[Li[q]];
    }
    // flop count of the numeric factorisation, SURVEY 8(d): sum_j (c_j^2 + 2 c_j), c_j = nnz(L(:,j))
    double flops() const { double f = 0; for (size_t j = 0; j + 1 < Lp.size(); j++) { double c = Lp[j + 1] - Lp[j]; f += c * c + 2 * c; } return f; }
};

// Exact minimum-degree ordering on the pattern of A + A^T (A: upper triangle).  Elimination graph
// with explicit adjacency sets: O(fill) memory, meant for the oracle's small/medium test problems.
inline IVec min_degree_ordering(const Csc& A) {
    const int n = A.rows;
    std::vector<std::set<int>> adj(n);
    for (int j = 0; j < n; j++) for (int q = A.p[j]; q < A.p[j + 1]; q++) { int i = A.i[q]; if (i != j) { adj[i].insert(j); adj[j].insert(i); } }
    std::set<std::pair<int, int>> heap;
    for (int v = 0; v < n; v++) heap.insert({(int)adj[v].size(), v});
    IVec perm; perm.reserve(n);
    std::vector<char> done(n, 0);
    while (!heap.empty()) {
        auto [deg, v] = *heap.begin(); heap.erase(heap.begin());
        (void)deg;
        perm.push_back(v); done[v] = 1;
        std::vector<int> nb(adj[v].begin(), adj[v].end());
        for (int u : nb) { heap.erase({(int)adj[u].size(), u}); adj[u].erase(v); }
        for (size_t a = 0; a < nb.size(); a++) for (size_t b = a + 1; b < nb.size(); b++) { adj[nb[a]].insert(nb[b]); adj[nb[b]].insert(nb[a]); }
        for (int u : nb) heap.insert({(int)adj[u].size(), u});
        adj[v].clear();
    }
    return perm;
}

struct SparseMatrices;

// sparse/kkt.hpp:31-250 with KKTImpl<FULL> (kkt_full.hpp)
struct SparseKKTFull : KKTBackend {
    const SparseMatrices& S;
    double m_delta = 0;   // reference leaves this uninitialised in the ctor (kkt.hpp:36,64); only dummy values depend on it
    Vec z_reg_inv, work_z, rhs, rhs_perm, P_diagonal;
    IVec perm, perm_inv;         // ordering.P / P_inv
    Csc PKPt;                    // permuted KKT, upper triangular
    IVec PKi;                    // KKT nz index -> PKPt nz index
    IVec P_to_Ki, AT_to_Ki, GT_to_Ki;
    SparseLDLt ldlt;

    SparseKKTFull(const SparseMatrices& S_, const IVec* user_perm);
    Csc create_kkt_matrix();
    static IVec permute_symmetric(const Csc& A, Csc& C, const IVec& inv);
    void update_data(int options) override;
    bool factor(double delta, const double* x_reg, const double* z_reg) override;
    void solve(const double* rx, const double* ry, const double* rz, double* lx, double* ly, double* lz) override;
    void eval_P_x(double alpha, const double* x, double* z) override;
    void eval_A(double an, double at, const double* xn, const double* xt, double* zn, double* zt) override;
    void eval_G(double an, double at, const double* xn, const double* xt, double* zn, double* zt) override;
};

inline std::unique_ptr<KKTBackend> make_multistage_backend(const SparseMatrices& S);  // oracle_multistage.hpp
inline std::unique_ptr<KKTBackend> make_cond_backend(const SparseMatrices& S, int mode, const IVec* user_perm);  // oracle_sparse_cond.hpp

struct SparseMatrices : QPMatrices {
    int n = 0, p = 0, m = 0;
    Csc P, AT, GT;  // P: upper triangle CSC; AT: n x p; GT: n x m
    IVec user_perm; // optional KKT ordering supplied by the caller (size n+p+m)
    int kkt_solver_hint = 1;   // Settings::kkt_solver at setup (lets a foreign backend factory pick its constructor)
    std::unique_ptr<KKTBackend> (*backend_factory)(const SparseMatrices&, void*) = nullptr;
    void* backend_factory_arg = nullptr;

    void set_G_row_zero(int row) override { for (int q = GT.p[row]; q < GT.p[row + 1]; q++) GT.x[q] = 0; }  // sparse/data.hpp:212-216
    void sym_col_absmax(Vec& v) const {
        for (int j = 0; j < n; j++) for (int q = P.p[j]; q < P.p[j + 1]; q++) {
            const int i = P.i[q]; const double a = std::fabs(P.x[q]);
            v[j] = std::max(v[j], a); if (i != j) v[i] = std::max(v[i], a);
        }
    }
    void kkt_col_norms(const Vec& xbs, Vec& nrm) override {  // sparse/preconditioner.hpp:85-130
        std::fill(nrm.begin(), nrm.end(), 0.0);
        for (int j = 0; j < n; j++) {
            for (int q = P.p[j]; q < P.p[j + 1]; q++) {
                const int i = P.i[q]; const double a = std::fabs(P.x[q]);
                nrm[j] = std::max(nrm[j], a); if (i != j) nrm[i] = std::max(nrm[i], a);
            }
            nrm[j] = std::max(nrm[j], xbs[j]);
        }
        for (int j = 0; j < p; j++) for (int q = AT.p[j]; q < AT.p[j + 1]; q++) { const double a = std::fabs(AT.x[q]); nrm[AT.i[q]] = std::max(nrm[AT.i[q]], a); nrm[n + j] = std::max(nrm[n + j], a); }
        for (int j = 0; j < m; j++) for (int q = GT.p[j]; q < GT.p[j + 1]; q++) { const double a = std::fabs(GT.x[q]); nrm[GT.i[q]] = std::max(nrm[GT.i[q]], a); nrm[n + p + j] = std::max(nrm[n + p + j], a); }
    }
    static void pre_post(Csc& A, const double* dr, const double* dc) {  // utils.hpp:171-201 (pre then post)
        for (int j = 0; j < A.cols; j++) for (int q = A.p[j]; q < A.p[j + 1]; q++) A.x[q] *= dr[A.i[q]];
        for (int j = 0; j < A.cols; j++) for (int q = A.p[j]; q < A.p[j + 1]; q++) A.x[q] *= dc[j];
    }
    void scale_sym(const double* d) override { pre_post(P, d, d); pre_post(AT, d, d + n); pre_post(GT, d, d + n + p); }
    double cost_norm_mean() override { Vec v(n, 0.0); sym_col_absmax(v); double s = 0; for (double e : v) s += e; return s / double(n); }
    void scale_P(double g) override { for (double& e : P.x) e *= g; }
    void extract_P_diag(Vec& dg) override { for (int j = 0; j < n; j++) for (int q = P.p[j]; q < P.p[j + 1]; q++) if (P.i[q] == j) dg[j] = P.x[q]; }
    std::unique_ptr<KKTBackend> make_backend(int kkt_solver) override {
        if (backend_factory) return backend_factory(*this, backend_factory_arg);
        if (kkt_solver == 5) return make_multistage_backend(*this);      // KKTSolver::sparse_multistage
        if (kkt_solver >= 2 && kkt_solver <= 4)                          // sparse_ldlt_eq_cond / _ineq_cond / _cond (kkt_system.hpp:476-489)
            return make_cond_backend(*this, kkt_solver - 1, user_perm.empty() ? nullptr : &user_perm);
        return std::make_unique<SparseKKTFull>(*this, user_perm.empty() ? nullptr : &user_perm);
    }
};

inline SparseKKTFull::SparseKKTFull(const SparseMatrices& S_, const IVec* user_perm) : S(S_) {  // kkt.hpp:51-70
    const int nk = S.n + S.p + S.m;
    z_reg_inv.assign(S.m, 0); work_z.assign(S.m, 0); rhs.assign(nk, 0); rhs_perm.assign(nk, 0);
    P_to_Ki.assign(S.P.nnz(), 0); P_diagonal.assign(S.n, 0); AT_to_Ki.assign(S.AT.nnz(), 0); GT_to_Ki.assign(S.GT.nnz(), 0);
    m_delta = 1.0;
    Csc K = create_kkt_matrix();
    perm = user_perm ? *user_perm : min_degree_ordering(K);
    perm_inv.assign(nk, 0);
    for (int i = 0; i < nk; i++) perm_inv[perm[i]] = i;
    PKi = permute_symmetric(K, PKPt, perm_inv);
    ldlt.symbolic(PKPt);
}

inline Csc SparseKKTFull::create_kkt_matrix() {  // kkt_full.hpp:39-170
    const int n = S.n, p = S.p, m = S.m, nk = n + p + m;
    Csc K; K.rows = K.cols = nk; K.p.assign(nk + 1, 0);
    int nz = 0, jk = 0;
    for (int j = 0; j < n; j++) {
        int cn = S.P.p[j + 1] - S.P.p[j];
        if (cn > 0) { if (S.P.i[S.P.p[j + 1] - 1] != j) cn++; } else cn++;
        nz += cn; K.p[++jk] = nz;
    }
    for (int j = 0; j < p; j++) { nz += S.AT.p[j + 1] - S.AT.p[j] + 1; K.p[++jk] = nz; }
    for (int j = 0; j < m; j++) { nz += S.GT.p[j + 1] - S.GT.p[j] + 1; K.p[++jk] = nz; }
    K.i.assign(nz, 0); K.x.assign(nz, 0);
    jk = 0;
    for (int j = 0; j < n; j++, jk++) {
        const int k0 = K.p[jk], cn = S.P.p[j + 1] - S.P.p[j], kcn = K.p[jk + 1] - k0;
        for (int t = 0; t < cn; t++) { K.i[k0 + t] = S.P.i[S.P.p[j] + t]; K.x[k0 + t] = S.P.x[S.P.p[j] + t]; P_to_Ki[S.P.p[j] + t] = k0 + t; }
        if (kcn > cn) { K.i[k0 + kcn - 1] = jk; K.x[k0 + kcn - 1] = 1.0; }
        else { P_diagonal[j] = S.P.x[S.P.p[j + 1] - 1]; K.x[k0 + kcn - 1] += 1.0; }
    }
    for (int j = 0; j < p; j++, jk++) {
        const int k0 = K.p[jk], cn = S.AT.p[j + 1] - S.AT.p[j];
        for (int t = 0; t < cn; t++) { K.i[k0 + t] = S.AT.i[S.AT.p[j] + t]; K.x[k0 + t] = S.AT.x[S.AT.p[j] + t]; AT_to_Ki[S.AT.p[j] + t] = k0 + t; }
        K.i[k0 + cn] = jk; K.x[k0 + cn] = -m_delta;
    }
    for (int j = 0; j < m; j++, jk++) {
        const int k0 = K.p[jk], cn = S.GT.p[j + 1] - S.GT.p[j];
        for (int t = 0; t < cn; t++) { K.i[k0 + t] = S.GT.i[S.GT.p[j] + t]; K.x[k0 + t] = S.GT.x[S.GT.p[j] + t]; GT_to_Ki[S.GT.p[j] + t] = k0 + t; }
        K.i[k0 + cn] = jk; K.x[k0 + cn] = -1.0 - m_delta;
    }
    return K;
}

// utils.hpp:31-128: C = upper(P A P^T) with sorted rows; returns map A nz -> C nz.
inline IVec SparseKKTFull::permute_symmetric(const Csc& A, Csc& C, const IVec& inv) {
    const int n = A.rows;
    IVec w(n, 0);
    for (int j = 0; j < n; j++) { const int j2 = inv[j];
        for (int q = A.p[j]; q < A.p[j + 1]; q++) { const int i = A.i[q]; if (i > j) continue; const int i2 = inv[i]; w[std::min(i2, j2)]++; } }
    Csc CT; CT.rows = CT.cols = n; CT.p.assign(n + 1, 0);
    int sum = 0;
    for (int i = 0; i < n; i++) { CT.p[i] = sum; sum += w[i]; w[i] = CT.p[i]; }
    CT.p[n] = sum; CT.i.assign(sum, 0); CT.x.assign(sum, 0);
    IVec CTi_to_Ai(sum);
    for (int j = 0; j < n; j++) { const int j2 = inv[j];
        for (int q = A.p[j]; q < A.p[j + 1]; q++) { const int i = A.i[q]; if (i > j) continue; const int i2 = inv[i];
            const int t = w[std::min(i2, j2)]++; CT.i[t] = std::max(i2, j2); CT.x[t] = A.x[q]; CTi_to_Ai[t] = q; } }
    C.rows = C.cols = n; C.p.assign(n + 1, 0);
    IVec cnt(n, 0);
    for (int t = 0; t < sum; t++) cnt[CT.i[t]]++;
    int s2 = 0;
    for (int j = 0; j < n; j++) { C.p[j] = s2; w[j] = s2; s2 += cnt[j]; }
    C.p[n] = s2; C.i.assign(s2, 0); C.x.assign(s2, 0);
    IVec Ai_to_Ci(A.nnz(), -1);
    for (int j = 0; j < n; j++) for (int t = CT.p[j]; t < CT.p[j + 1]; t++) { const int i = CT.i[t]; const int q = w[i]++; C.i[q] = j; C.x[q] = CT.x[t]; Ai_to_Ci[CTi_to_Ai[t]] = q; }
    return Ai_to_Ci;
}

inline void SparseKKTFull::update_data(int options) {  // kkt_full.hpp:212-251
    if (options & UPDATE_P) for (int j = 0; j < S.n; j++) for (int q = S.P.p[j]; q < S.P.p[j + 1]; q++) {
        PKPt.x[PKi[P_to_Ki[q]]] = S.P.x[q]; if (S.P.i[q] == j) P_diagonal[j] = S.P.x[q]; }
    if (options & UPDATE_A) for (int q = 0; q < S.AT.nnz(); q++) PKPt.x[PKi[AT_to_Ki[q]]] = S.AT.x[q];
    if (options & UPDATE_G) for (int q = 0; q < S.GT.nnz(); q++) PKPt.x[PKi[GT_to_Ki[q]]] = S.GT.x[q];
}

inline bool SparseKKTFull::factor(double delta, const double* x_reg, const double* z_reg) {  // kkt.hpp:83-105 + kkt_full.hpp:172-210
    const int n = S.n, p = S.p, m = S.m;
    m_delta = delta;
    for (int i = 0; i < m; i++) z_reg_inv[i] = 1.0 / z_reg[i];
    for (int c = 0; c < n; c++) PKPt.x[PKPt.p[perm_inv[c] + 1] - 1] = P_diagonal[c] + x_reg[c];
    for (int c = n; c < n + p; c++) PKPt.x[PKPt.p[perm_inv[c] + 1] - 1] = -m_delta;
    for (int c = n + p, k = 0; c < n + p + m; c++, k++) PKPt.x[PKPt.p[perm_inv[c] + 1] - 1] = -z_reg[k];
    return ldlt.numeric(PKPt) == PKPt.cols;
}

inline void SparseKKTFull::solve(const double* rx, const double* ry, const double* rz, double* lx, double* ly, double* lz) {  // kkt.hpp:107-176 (FULL)
    const int n = S.n, p = S.p, m = S.m, nk = n + p + m;
    for (int i = 0; i < n; i++) rhs[i] = rx[i];
    for (int i = 0; i < p; i++) rhs[n + i] = ry[i];
    for (int i = 0; i < m; i++) rhs[n + p + i] = rz[i];
    for (int j = 0; j < nk; j++) rhs_perm[j] = rhs[perm[j]];
    ldlt.solve_inplace(rhs_perm.data());
    for (int j = 0; j < nk; j++) rhs[perm[j]] = rhs_perm[j];
    for (int i = 0; i < n; i++) lx[i] = rhs[i];
    for (int i = 0; i < p; i++) ly[i] = rhs[n + i];
    for (int i = 0; i < m; i++) lz[i] = rhs[n + p + i];
}

inline void csc_mv_nt(const Csc& AT, double an, double at, const double* xn, const double* xt, double* zn, double* zt) {
    // zn = an * AT^T xn (per column dot), zt = at * AT xt (column axpy)
    for (int i = 0; i < AT.rows; i++) zt[i] = 0;
    for (int j = 0; j < AT.cols; j++) {
        double s = 0; const double w = at * xt[j];
        for (int q = AT.p[j]; q < AT.p[j + 1]; q++) { s += AT.x[q] * xn[AT.i[q]]; zt[AT.i[q]] += AT.x[q] * w; }
        zn[j] = an * s;
    }
}
inline void SparseKKTFull::eval_P_x(double alpha, const double* x, double* z) {  // kkt.hpp:179-185
    for (int i = 0; i < S.n; i++) z[i] = 0;
    for (int j = 0; j < S.n; j++) for (int q = S.P.p[j]; q < S.P.p[j + 1]; q++) {
        const int i = S.P.i[q]; z[i] += alpha * S.P.x[q] * x[j]; if (i != j) z[j] += alpha * S.P.x[q] * x[i]; }
}
inline void SparseKKTFull::eval_A(double an, double at, const double* xn, const double* xt, double* zn, double* zt) { csc_mv_nt(S.AT, an, at, xn, xt, zn, zt); }
inline void SparseKKTFull::eval_G(double an, double at, const double* xn, const double* xt, double* zn, double* zt) { csc_mv_nt(S.GT, an, at, xn, xt, zn, zt); }

}  // namespace oracle
