// oracle/oracle_sparse_cond.hpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement of the three CONDENSED modes of PIQP's general sparse KKT backend (v0.6.2):
//   KKTImpl<KKT_EQ_ELIMINATED>    include/piqp/sparse/kkt_eq_eliminated.hpp:34-252     (sparse_ldlt_eq_cond,   n_kkt = n + m)
//   KKTImpl<KKT_INEQ_ELIMINATED>  include/piqp/sparse/kkt_ineq_eliminated.hpp:34-253   (sparse_ldlt_ineq_cond, n_kkt = n + p)
//   KKTImpl<KKT_ALL_ELIMINATED>   include/piqp/sparse/kkt_all_eliminated.hpp:36-224    (sparse_ldlt_cond,      n_kkt = n)
// driven by sparse::KKT<Mode> (include/piqp/sparse/kkt.hpp:51-176): rhs condensation before and dual recovery after the
// LDL^T solve.  The top-left block is  P + diag(x_reg) [+ delta^-1 A^T A] [+ G^T Z^-1 G]  on the structural union pattern.
//
// Third-party arithmetic: the reference forms the INITIAL A^T A with Eigen's sparse product (kkt_eq_eliminated.hpp:37,
// kkt_all_eliminated.hpp:43); later updates use its own Gustavson loop (update_AT_A).  The oracle uses the Gustavson
// loop for both (same sums, possibly different rounding of the initial product: parity unpinned at that level).
#pragma once
#include "oracle_sparse.hpp"

namespace oracle {

struct SparseKKTCond : KKTBackend {
    const SparseMatrices& S;
    const bool elim_eq, elim_ineq;      // KKTMode bits (kkt_fwd.hpp:15-21)
    double m_delta = 1.0;               // uninitialised in the reference ctor; only dummy values depend on it
    Vec z_reg_inv, work_z, rhs, rhs_perm, tmp_scatter;
    IVec perm, perm_inv;
    Csc PKPt; IVec PKi;
    Csc A, G;                           // A = AT^T (p x n), G = GT^T (m x n)
    Csc AT_A, GT_G;                     // upper(AT * A), upper(GT * Z^-1 * G)
    IVec P_to_Ki, AT_A_to_Ki, GT_G_to_Ki, AT_to_Ki, GT_to_Ki;
    SparseLDLt ldlt;
    int nk = 0;

    static Csc transpose(const Csc& M) {   // utils.hpp:131-163 (transpose_no_allocation computes the same result)
        Csc T; T.rows = M.cols; T.cols = M.rows; T.p.assign(M.rows + 1, 0); T.i.assign(M.nnz(), 0); T.x.assign(M.nnz(), 0);
        for (int q = 0; q < M.nnz(); q++) T.p[M.i[q] + 1]++;
        for (int r = 0; r < M.rows; r++) T.p[r + 1] += T.p[r];
        IVec w(T.p.begin(), T.p.end() - 1);
        for (int j = 0; j < M.cols; j++) for (int q = M.p[j]; q < M.p[j + 1]; q++) { const int t = w[M.i[q]]++; T.i[t] = j; T.x[t] = M.x[q]; }
        return T;
    }
    // structural pattern of upper(MT * M), MT: n x r, M = MT^T: r x n
    static Csc gram_pattern(const Csc& MT, const Csc& M) {
        const int n = MT.rows;
        Csc C; C.rows = C.cols = n; C.p.assign(n + 1, 0);
        IVec mark(n, -1);
        for (int j = 0; j < n; j++) {
            IVec rows;
            for (int a = M.p[j]; a < M.p[j + 1]; a++) { const int k = M.i[a];
                for (int t = MT.p[k]; t < MT.p[k + 1]; t++) { const int i = MT.i[t]; if (i > j) continue; if (mark[i] != j) { mark[i] = j; rows.push_back(i); } } }
            std::sort(rows.begin(), rows.end());
            for (int i : rows) C.i.push_back(i);
            C.p[j + 1] = (int)C.i.size();
        }
        C.x.assign(C.i.size(), 0.0);
        return C;
    }
    // update_AT_A / update_GT_W_delta_inv_G (kkt_all_eliminated.hpp:178-220): Gustavson product with a dense scatter vector
    void update_gram(const Csc& M, const Csc& MT, Csc& C, const double* z_reg) {
        for (int j = 0; j < M.cols; j++) {
            for (int a = M.p[j]; a < M.p[j + 1]; a++) { const int k = M.i[a];
                for (int t = MT.p[k]; t < MT.p[k + 1]; t++) { const int i = MT.i[t]; if (i > j) continue;
                    if (z_reg) tmp_scatter[i] += M.x[a] * MT.x[t] / z_reg[k]; else tmp_scatter[i] += M.x[a] * MT.x[t]; } }
            for (int q = C.p[j]; q < C.p[j + 1]; q++) { C.x[q] = tmp_scatter[C.i[q]]; tmp_scatter[C.i[q]] = 0; }
        }
    }

    SparseKKTCond(const SparseMatrices& S_, int mode, const IVec* user_perm) : S(S_), elim_eq(mode & 1), elim_ineq(mode & 2) {  // kkt.hpp:51-70
        const int n = S.n, p = S.p, m = S.m;
        nk = n + (elim_eq ? 0 : p) + (elim_ineq ? 0 : m);
        z_reg_inv.assign(m, 0); work_z.assign(m, 0); rhs.assign(nk, 0); rhs_perm.assign(nk, 0);
        tmp_scatter.assign(std::max(1, n), 0.0);
        // init_workspace
        if (elim_eq) { A = transpose(S.AT); AT_A = gram_pattern(S.AT, A); update_gram(A, S.AT, AT_A, nullptr); }
        if (elim_ineq) { G = transpose(S.GT); GT_G = gram_pattern(S.GT, G); }     // values are set at every factor
        P_to_Ki.assign(S.P.nnz(), 0); AT_A_to_Ki.assign(AT_A.nnz(), 0); GT_G_to_Ki.assign(GT_G.nnz(), 0);
        AT_to_Ki.assign(S.AT.nnz(), 0); GT_to_Ki.assign(S.GT.nnz(), 0);
        Csc K = create_kkt_matrix();
        perm = user_perm ? *user_perm : min_degree_ordering(K);
        perm_inv.assign(nk, 0);
        for (int i = 0; i < nk; i++) perm_inv[perm[i]] = i;
        PKi = SparseKKTFull::permute_symmetric(K, PKPt, perm_inv);
        ldlt.symbolic(PKPt);
    }

    // pattern = union of P_utri, I, AT_A, GT_G in the top-left block; then [AT; -delta] and/or [GT; -Z] columns.
    // Values only matter as structure (every factor rewrites them).
    Csc create_kkt_matrix() {
        const int n = S.n, p = S.p, m = S.m;
        Csc K; K.rows = K.cols = nk; K.p.assign(nk + 1, 0);
        for (int j = 0; j < n; j++) {
            int a = S.P.p[j], a1 = S.P.p[j + 1];
            int b = elim_eq ? AT_A.p[j] : 0, b1 = elim_eq ? AT_A.p[j + 1] : 0;
            int c = elim_ineq ? GT_G.p[j] : 0, c1 = elim_ineq ? GT_G.p[j + 1] : 0;
            bool diag_done = false;
            while (true) {
                int r = j + 1;     // sentinel: larger than any upper row
                if (a < a1) r = std::min(r, S.P.i[a]);
                if (b < b1) r = std::min(r, AT_A.i[b]);
                if (c < c1) r = std::min(r, GT_G.i[c]);
                if (!diag_done) r = std::min(r, j);
                if (r > j) break;
                const int e = (int)K.i.size();
                K.i.push_back(r); K.x.push_back(1.0);
                if (a < a1 && S.P.i[a] == r) P_to_Ki[a++] = e;
                if (b < b1 && AT_A.i[b] == r) AT_A_to_Ki[b++] = e;
                if (c < c1 && GT_G.i[c] == r) GT_G_to_Ki[c++] = e;
                if (r == j) diag_done = true;
            }
            K.p[j + 1] = (int)K.i.size();
        }
        int jk = n;
        if (!elim_eq) for (int j = 0; j < p; j++, jk++) {
            for (int q = S.AT.p[j]; q < S.AT.p[j + 1]; q++) { AT_to_Ki[q] = (int)K.i.size(); K.i.push_back(S.AT.i[q]); K.x.push_back(S.AT.x[q]); }
            K.i.push_back(jk); K.x.push_back(-m_delta);
            K.p[jk + 1] = (int)K.i.size();
        }
        if (!elim_ineq) for (int j = 0; j < m; j++, jk++) {
            for (int q = S.GT.p[j]; q < S.GT.p[j + 1]; q++) { GT_to_Ki[q] = (int)K.i.size(); K.i.push_back(S.GT.i[q]); K.x.push_back(S.GT.x[q]); }
            K.i.push_back(jk); K.x.push_back(-1.0 - m_delta);
            K.p[jk + 1] = (int)K.i.size();
        }
        return K;
    }

    void update_data(int options) override {   // update_data_impl of the three files
        if ((options & UPDATE_A) && elim_eq) { A = transpose(S.AT); update_gram(A, S.AT, AT_A, nullptr); }
        if ((options & UPDATE_G) && elim_ineq) G = transpose(S.GT);
    }

    bool factor(double delta, const double* x_reg, const double* z_reg) override {   // kkt.hpp:83-105
        const int n = S.n, p = S.p, m = S.m;
        m_delta = delta;
        for (int i = 0; i < m; i++) z_reg_inv[i] = 1.0 / z_reg[i];
        auto diag = [&](int col) -> double& { return PKPt.x[PKPt.p[perm_inv[col] + 1] - 1]; };
        // update_kkt_cost_scalings
        std::fill(PKPt.x.begin(), PKPt.x.end(), 0.0);
        for (int q = 0; q < S.P.nnz(); q++) PKPt.x[PKi[P_to_Ki[q]]] += S.P.x[q];
        for (int c = 0; c < n; c++) diag(c) += x_reg[c];
        // update_kkt_equality_scalings
        int col = n;
        if (elim_eq) { const double dinv = 1.0 / m_delta; for (int q = 0; q < AT_A.nnz(); q++) PKPt.x[PKi[AT_A_to_Ki[q]]] += dinv * AT_A.x[q]; }
        else { for (int q = 0; q < S.AT.nnz(); q++) PKPt.x[PKi[AT_to_Ki[q]]] = S.AT.x[q]; for (int k = 0; k < p; k++, col++) diag(col) = -m_delta; }
        // update_kkt_inequality_scaling
        if (elim_ineq) { update_gram(G, S.GT, GT_G, z_reg); for (int q = 0; q < GT_G.nnz(); q++) PKPt.x[PKi[GT_G_to_Ki[q]]] += GT_G.x[q]; }
        else { for (int q = 0; q < S.GT.nnz(); q++) PKPt.x[PKi[GT_to_Ki[q]]] = S.GT.x[q]; for (int k = 0; k < m; k++, col++) diag(col) = -z_reg[k]; }
        return ldlt.numeric(PKPt) == PKPt.cols;
    }

    void solve(const double* rx, const double* ry, const double* rz, double* lx, double* ly, double* lz) override {   // kkt.hpp:107-176
        const int n = S.n, p = S.p, m = S.m;
        const double dinv = 1.0 / m_delta;
        for (int i = 0; i < n; i++) rhs[i] = rx[i];
        if (elim_ineq) {
            for (int k = 0; k < m; k++) work_z[k] = z_reg_inv[k] * rz[k];
            for (int k = 0; k < m; k++) for (int q = S.GT.p[k]; q < S.GT.p[k + 1]; q++) rhs[S.GT.i[q]] += S.GT.x[q] * work_z[k];
        }
        if (elim_eq) for (int k = 0; k < p; k++) { const double w = dinv * ry[k]; for (int q = S.AT.p[k]; q < S.AT.p[k + 1]; q++) rhs[S.AT.i[q]] += S.AT.x[q] * w; }
        int o = n;
        if (!elim_eq) { for (int k = 0; k < p; k++) rhs[o + k] = ry[k]; o += p; }
        if (!elim_ineq) { for (int k = 0; k < m; k++) rhs[o + k] = rz[k]; }
        for (int j = 0; j < nk; j++) rhs_perm[j] = rhs[perm[j]];
        ldlt.solve_inplace(rhs_perm.data());
        for (int j = 0; j < nk; j++) rhs[perm[j]] = rhs_perm[j];
        for (int i = 0; i < n; i++) lx[i] = rhs[i];
        o = n;
        if (elim_eq) {
            for (int k = 0; k < p; k++) { double s = 0; for (int q = S.AT.p[k]; q < S.AT.p[k + 1]; q++) s += S.AT.x[q] * lx[S.AT.i[q]]; ly[k] = dinv * s; ly[k] -= dinv * ry[k]; }
        } else { for (int k = 0; k < p; k++) ly[k] = rhs[o + k]; o += p; }
        if (elim_ineq) {
            for (int k = 0; k < m; k++) { double s = 0; for (int q = S.GT.p[k]; q < S.GT.p[k + 1]; q++) s += S.GT.x[q] * lx[S.GT.i[q]]; lz[k] = s; lz[k] -= rz[k]; lz[k] *= z_reg_inv[k]; }
        } else { for (int k = 0; k < m; k++) lz[k] = rhs[o + k]; }
    }

    void eval_P_x(double alpha, const double* x, double* z) override {
        for (int i = 0; i < S.n; i++) z[i] = 0;
        for (int j = 0; j < S.n; j++) for (int q = S.P.p[j]; q < S.P.p[j + 1]; q++) {
            const int i = S.P.i[q]; z[i] += alpha * S.P.x[q] * x[j]; if (i != j) z[j] += alpha * S.P.x[q] * x[i]; }
    }
    void eval_A(double an, double at, const double* xn, const double* xt, double* zn, double* zt) override { csc_mv_nt(S.AT, an, at, xn, xt, zn, zt); }
    void eval_G(double an, double at, const double* xn, const double* xt, double* zn, double* zt) override { csc_mv_nt(S.GT, an, at, xn, xt, zn, zt); }
};

inline std::unique_ptr<KKTBackend> make_cond_backend(const SparseMatrices& S, int mode, const IVec* user_perm) {
    return std::make_unique<SparseKKTCond>(S, mode, user_perm);
}

}  // namespace oracle
