// oracle/oracle_multistage.hpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement of PIQP's block-tridiagonal-arrow backend `sparse::MultistageKKT`
// (include/piqp/sparse/multistage_kkt.hpp):
//   extract_arrow_structure   :420-597   greedy tri-diagonal-vs-arrow flop model, block merge pass
//   utri_to_kkt / block layout:599-670   D_i (d_i x d_i), B_i (o_i x d_i), E_i (w x d_i), D_N (w x w)
//   block_syrk_ln             :833-994   AtA / GtG in block form
//   construct_kkt_fac         :1008-1219 kkt = P + AtA/delta + GtG + diag(x_reg)
//   factor_kkt                :1253-1352 block Cholesky recursion
//   solve_llt_in_place        :1709-1816 block forward / backward substitution
//   solve / eval_*            :221-383
// The arithmetic of BLASFEO (third party, not under /root/reference, unpinned) is replaced by plain loops on
// column-major dense blocks; A and G contributions are accumulated row by row straight into the block storage
// instead of through the reference's row-permuted BlockMat (same sums, different order).  Like the reference this
// backend never reports a factorisation failure (:218).
#pragma once
#include "oracle_sparse.hpp"
#include <cassert>

namespace oracle {

struct BlockInfo { int start, diag, off; };   // blocksparse/block_info.hpp:18

// multistage_kkt.hpp:420-597 on the structural pattern of C = P_ltri + I + AtA_lower + GtG_lower
inline std::vector<BlockInfo> extract_arrow_structure(int n, const Csc& P, const Csc& AT, const Csc& GT) {
    // rows[i] = sorted structural column indices j >= i of row i of the upper triangle of C^T (== column i of lower C)
    std::vector<std::vector<int>> up(n);
    for (int j = 0; j < n; j++) for (int q = P.p[j]; q < P.p[j + 1]; q++) up[P.i[q]].push_back(j);   // P_utri(i,j), i <= j
    for (int i = 0; i < n; i++) up[i].push_back(i);
    auto add_rows = [&](const Csc& MT) {
        for (int r = 0; r < MT.cols; r++)
            for (int a = MT.p[r]; a < MT.p[r + 1]; a++)
                for (int b = a; b < MT.p[r + 1]; b++) up[MT.i[a]].push_back(MT.i[b]);   // indices sorted within a column
    };
    add_rows(AT); add_rows(GT);
    for (auto& v : up) { std::sort(v.begin(), v.end()); v.erase(std::unique(v.begin(), v.end()), v.end()); }

    typedef unsigned long long usize;
    auto f_gemm = [](usize m, usize nn, usize k) { return 2 * m * nn * k; };
    auto f_trsm = [](usize m, usize nn) { return m * m * nn; };
    auto f_syrk = [](usize nn, usize k) { return nn * nn * k; };
    auto f_potrf = [](usize nn) { return nn * nn * nn / 3; };

    struct Info { int prev_diag = 0, start = 0, diag = 0, off = 0, arrow = 0; };
    Info cur;
    usize fl_tri = 0, fl_arrow_nosyrk = 0, fl_arrow_syrk = 0;
    std::vector<BlockInfo> blocks;

    auto next_structure = [&](int row, const Info& ci) {
        Info ni = ci;
        if (row >= n) return ni;
        for (int col : up[row]) {
            if (col >= ni.start && col + ni.arrow < n) {
                const int cur_size = ni.diag + ni.off;
                const int new_size = std::max(col - ni.start + 1, cur_size);
                const int max_diag = row - ni.start + 1;
                const int new_min_diag = std::max(ni.diag, (new_size + 1) / 2);
                const int new_diag = std::max(new_min_diag, max_diag);
                const int new_off = new_size - new_diag;
                const int remaining = n - ni.start - ni.diag - ni.off;
                const int new_arrow = std::min(std::max(ni.arrow, n - col), remaining);

                usize tri_new = fl_tri;
                tri_new += f_syrk((usize)new_diag, (usize)ni.prev_diag);
                tri_new += f_potrf((usize)new_diag);
                tri_new += f_trsm((usize)new_diag, (usize)new_off);

                const usize aw = (usize)((ni.arrow + 3) / 4) * 4, naw = (usize)((new_arrow + 3) / 4) * 4;
                usize arrow = aw * fl_arrow_nosyrk + aw * aw * fl_arrow_syrk + f_potrf(aw);
                usize arrow_new = naw * fl_arrow_nosyrk + naw * naw * fl_arrow_syrk;
                arrow_new += f_gemm(naw, (usize)ni.prev_diag, (usize)new_diag);
                arrow_new += f_trsm((usize)new_diag, naw);
                arrow_new += f_syrk(naw, (usize)new_diag);
                arrow_new += f_potrf(naw);
                if (tri_new - fl_tri <= arrow_new - arrow) { ni.diag = new_diag; ni.off = new_off; }
                else ni.arrow = new_arrow;
            }
        }
        return ni;
    };

    for (int i = 0; i < n; i++) {
        cur = next_structure(i, cur);
        if (i + 1 >= cur.start + cur.diag) {
            const bool hit_ratio = cur.diag >= 2 * cur.off;
            const bool at_end = i + 1 >= n - cur.arrow;
            auto next_grows = [&]() {
                if (i >= n) return false;
                Info nx = next_structure(i + 1, cur);
                return nx.diag + nx.off > cur.diag + cur.off;
            };
            if (hit_ratio || at_end || next_grows()) {
                blocks.push_back({cur.start, cur.diag, cur.off});
                fl_tri += f_syrk((usize)cur.diag, (usize)(cur.prev_diag + 1));
                fl_tri += f_potrf((usize)cur.diag);
                fl_tri += f_trsm((usize)cur.diag, (usize)cur.off);
                fl_arrow_nosyrk += f_gemm(1, (usize)cur.prev_diag, (usize)cur.diag);
                fl_arrow_nosyrk += f_trsm((usize)cur.diag, 1);
                fl_arrow_syrk += f_syrk(1, (usize)cur.diag);
                cur.start += cur.diag; cur.prev_diag = cur.diag; cur.diag = cur.off; cur.off = 0;
            }
            if (at_end && cur.diag > 0) {
                blocks.push_back({cur.start, cur.diag, cur.off});
                cur.start += cur.diag; cur.prev_diag = cur.diag; cur.diag = cur.off; cur.off = 0;
            }
            if (at_end) break;
        }
    }
    // merge blocks which are split in two (:569-579)
    for (size_t i = 0; i + 1 < blocks.size(); i++) {
        if (blocks[i].off == blocks[i + 1].diag && blocks[i + 1].off == 0) {
            blocks[i].diag += blocks[i].off; blocks[i].off = 0;
            blocks.erase(blocks.begin() + (long)i + 1);
        }
    }
    blocks.push_back({cur.start, cur.arrow, 0});   // arrow corner block
    return blocks;
}

// dense column-major block
struct DBlock { int r = 0, c = 0; Vec v; void init(int r_, int c_) { r = r_; c = c_; v.assign((size_t)r * c, 0.0); }
                double& at(int i, int j) { return v[i + (size_t)j * r]; } double at(int i, int j) const { return v[i + (size_t)j * r]; } };

// BlockKKT (blocksparse/block_kkt.hpp:22-34) with all blocks allocated
struct BlockKKTStore {
    std::vector<DBlock> D, B, E;
    void init(const std::vector<BlockInfo>& bi) {
        const size_t N = bi.size(); const int w = bi.back().diag;
        D.resize(N); B.resize(N >= 2 ? N - 2 : 0); E.resize(N - 1);
        for (size_t i = 0; i < N; i++) D[i].init(bi[i].diag, bi[i].diag);
        for (size_t i = 0; i + 2 < N; i++) B[i].init(bi[i].off, bi[i].diag);
        for (size_t i = 0; i + 1 < N; i++) E[i].init(w, bi[i].diag);
    }
    void zero() { for (auto& x : D) std::fill(x.v.begin(), x.v.end(), 0.0); for (auto& x : B) std::fill(x.v.begin(), x.v.end(), 0.0); for (auto& x : E) std::fill(x.v.begin(), x.v.end(), 0.0); }
};

struct MultistageKKT : KKTBackend {
    const SparseMatrices& S;
    std::vector<BlockInfo> bi;
    IVec blk_of;                 // variable -> block index
    BlockKKTStore Pb, AtAb, GtGb, fac;
    double m_delta = 1.0;
    Vec z_reg_inv, work_z, bx;

    explicit MultistageKKT(const SparseMatrices& S_) : S(S_) {   // :74-133
        bi = extract_arrow_structure(S.n, S.P, S.AT, S.GT);
        blk_of.assign(S.n, 0);
        for (size_t b = 0; b < bi.size(); b++) for (int k = 0; k < bi[b].diag; k++) blk_of[bi[b].start + k] = (int)b;
        Pb.init(bi); AtAb.init(bi); GtGb.init(bi); fac.init(bi);
        z_reg_inv.assign(S.m, 0); work_z.assign(S.m, 0); bx.assign(S.n, 0);
        load_P(); accumulate(S.AT, nullptr, AtAb);
    }
    // element (i >= j) of the condensed matrix -> its slot; returns nullptr if outside the block structure
    double* slot(BlockKKTStore& K, int i, int j) {
        const int N = (int)bi.size(), w = bi.back().diag, n = S.n;
        const int bj = blk_of[j], bi_ = blk_of[i];
        if (bi_ == bj) return &K.D[bj].at(i - bi[bj].start, j - bi[bj].start);
        if (w > 0 && i >= n - w) return &K.E[bj].at(i - (n - w), j - bi[bj].start);
        if (bi_ == bj + 1 && bj + 2 < N && i - bi[bi_].start < bi[bj].off) return &K.B[bj].at(i - bi[bi_].start, j - bi[bj].start);
        return nullptr;
    }
    void load_P() {   // utri_to_kkt :599-670
        Pb.zero();
        for (int j = 0; j < S.n; j++) for (int q = S.P.p[j]; q < S.P.p[j + 1]; q++) {
            double* s = slot(Pb, j, S.P.i[q]);   // P_utri(i,j) with i <= j is lower element (j, i)
            assert(s && "P entry outside the detected block structure");
            if (s) *s = S.P.x[q];
        }
    }
    // K += sum_r w_r * a_r a_r^T (lower part), rows a_r = columns of MT   (block_syrk_ln :833-994)
    void accumulate(const Csc& MT, const double* w, BlockKKTStore& K) {
        K.zero();
        for (int r = 0; r < MT.cols; r++) {
            const double wr = w ? w[r] : 1.0;
            for (int a = MT.p[r]; a < MT.p[r + 1]; a++)
                for (int b = MT.p[r]; b <= a; b++) {
                    double* s = slot(K, MT.i[a], MT.i[b]);
                    assert(s && "constraint row couples variables outside the detected block structure");
                    if (s) *s += (wr * MT.x[a]) * MT.x[b];
                }
        }
    }
    void update_data(int options) override {   // :140-178
        if (options & UPDATE_P) load_P();
        if (options & UPDATE_A) accumulate(S.AT, nullptr, AtAb);
    }

    bool factor(double delta, const double* x_reg, const double* z_reg) override {   // :180-219
        m_delta = delta;
        for (int i = 0; i < S.m; i++) z_reg_inv[i] = 1.0 / z_reg[i];
        accumulate(S.GT, z_reg_inv.data(), GtGb);
        const double dinv = 1.0 / m_delta;
        const size_t N = bi.size();
        auto comb = [&](DBlock& o, const DBlock& p, const DBlock& a, const DBlock& g) { for (size_t k = 0; k < o.v.size(); k++) { double v = p.v[k]; v += dinv * a.v[k]; v += g.v[k]; o.v[k] = v; } };
        for (size_t i = 0; i < N; i++) { comb(fac.D[i], Pb.D[i], AtAb.D[i], GtGb.D[i]); for (int k = 0; k < bi[i].diag; k++) fac.D[i].at(k, k) += x_reg[bi[i].start + k]; }
        for (size_t i = 0; i + 2 < N; i++) comb(fac.B[i], Pb.B[i], AtAb.B[i], GtGb.B[i]);
        for (size_t i = 0; i + 1 < N; i++) comb(fac.E[i], Pb.E[i], AtAb.E[i], GtGb.E[i]);
        factor_kkt();
        return true;   // the reference never signals failure here (:218)
    }

    // in-place lower Cholesky of a d x d block (dpotrf_l); no failure reporting, like BLASFEO
    static void potrf(DBlock& A) {
        const int d = A.r;
        for (int k = 0; k < d; k++) {
            double x = A.at(k, k);
            for (int j = 0; j < k; j++) x -= A.at(k, j) * A.at(k, j);
            x = std::sqrt(x);
            A.at(k, k) = x;
            for (int i = k + 1; i < d; i++) { double s = A.at(i, k); for (int j = 0; j < k; j++) s -= A.at(i, j) * A.at(k, j); A.at(i, k) = s / x; }
        }
    }
    // X <- X * L^{-T}  (dtrsm_rltn), X is m x d
    static void trsm_rltn(const DBlock& L, DBlock& X) {
        const int d = L.r;
        for (int i = 0; i < X.r; i++)
            for (int j = 0; j < d; j++) { double s = X.at(i, j); for (int k = 0; k < j; k++) s -= X.at(i, k) * L.at(j, k); X.at(i, j) = s / L.at(j, j); }
    }
    // C(lower, leading m x m) -= A * A^T, A is m x k
    static void syrk_sub(DBlock& C, const DBlock& A) {
        for (int j = 0; j < A.r; j++) for (int i = j; i < A.r; i++) { double s = 0; for (int k = 0; k < A.c; k++) s += A.at(i, k) * A.at(j, k); C.at(i, j) -= s; }
    }
    // C(m x n) -= A(m x k) * B(n x k)^T on the leading n columns of C
    static void gemm_nt_sub(DBlock& C, const DBlock& A, const DBlock& B) {
        for (int j = 0; j < B.r; j++) for (int i = 0; i < A.r; i++) { double s = 0; for (int k = 0; k < A.c; k++) s += A.at(i, k) * B.at(j, k); C.at(i, j) -= s; }
    }
    void factor_kkt() {   // :1253-1352
        const size_t N = bi.size(); const int w = bi.back().diag;
        for (size_t i = 0; i + 1 < N; i++) {
            if (i > 0 && bi[i - 1].off > 0) syrk_sub(fac.D[i], fac.B[i - 1]);              // L_i = chol(D_i - C_{i-1} C_{i-1}^T)
            potrf(fac.D[i]);
            if (i + 2 < N && bi[i].off > 0) trsm_rltn(fac.D[i], fac.B[i]);               // C_i = B_i L_i^{-T}
            if (w > 0) {
                if (i > 0 && bi[i - 1].off > 0) gemm_nt_sub(fac.E[i], fac.E[i - 1], fac.B[i - 1]);   // E_i -= F_{i-1} C_{i-1}^T
                trsm_rltn(fac.D[i], fac.E[i]);                                            // F_i = E_i L_i^{-T}
                syrk_sub(fac.D[N - 1], fac.E[i]);                                         // D_N -= F_i F_i^T
            }
        }
        if (w > 0) potrf(fac.D[N - 1]);
    }
    void solve_llt(double* x) {   // :1709-1816
        const size_t N = bi.size(); const int w = bi.back().diag, n = S.n;
        auto lsolve = [&](const DBlock& L, double* v) { for (int j = 0; j < L.r; j++) { double s = v[j]; for (int k = 0; k < j; k++) s -= L.at(j, k) * v[k]; v[j] = s / L.at(j, j); } };
        auto ltsolve = [&](const DBlock& L, double* v) { for (int j = L.r - 1; j >= 0; j--) { double s = v[j]; for (int k = j + 1; k < L.r; k++) s -= L.at(k, j) * v[k]; v[j] = s / L.at(j, j); } };
        for (size_t i = 0; i + 1 < N; i++) {
            double* xi = x + bi[i].start;
            if (i > 0 && bi[i - 1].off > 0) { const DBlock& C = fac.B[i - 1]; const double* xp = x + bi[i - 1].start;
                for (int r = 0; r < C.r; r++) { double s = 0; for (int k = 0; k < C.c; k++) s += C.at(r, k) * xp[k]; xi[r] -= s; } }
            lsolve(fac.D[i], xi);
        }
        if (w > 0) {
            double* xn = x + (n - w);
            for (size_t i = 0; i + 1 < N; i++) { const DBlock& F = fac.E[i]; const double* xi = x + bi[i].start;
                for (int r = 0; r < w; r++) { double s = 0; for (int k = 0; k < F.c; k++) s += F.at(r, k) * xi[k]; xn[r] -= s; } }
            lsolve(fac.D[N - 1], xn);
            ltsolve(fac.D[N - 1], xn);
        }
        for (size_t ii = N - 1; ii-- > 0;) {
            double* xi = x + bi[ii].start;
            if (ii + 2 < N && bi[ii].off > 0) { const DBlock& C = fac.B[ii]; const double* xq = x + bi[ii + 1].start;
                for (int k = 0; k < C.c; k++) { double s = 0; for (int r = 0; r < C.r; r++) s += C.at(r, k) * xq[r]; xi[k] -= s; } }
            if (w > 0) { const DBlock& F = fac.E[ii]; const double* xn = x + (n - w);
                for (int k = 0; k < F.c; k++) { double s = 0; for (int r = 0; r < w; r++) s += F.at(r, k) * xn[r]; xi[k] -= s; } }
            ltsolve(fac.D[ii], xi);
        }
    }
    void solve(const double* rx, const double* ry, const double* rz, double* lx, double* ly, double* lz) override {   // :221-288
        const int n = S.n, p = S.p, m = S.m;
        const double dinv = 1.0 / m_delta;
        for (int i = 0; i < m; i++) work_z[i] = z_reg_inv[i] * rz[i];
        for (int i = 0; i < n; i++) bx[i] = rx[i];
        for (int k = 0; k < m; k++) for (int q = S.GT.p[k]; q < S.GT.p[k + 1]; q++) bx[S.GT.i[q]] += S.GT.x[q] * work_z[k];
        for (int k = 0; k < p; k++) for (int q = S.AT.p[k]; q < S.AT.p[k + 1]; q++) bx[S.AT.i[q]] += dinv * S.AT.x[q] * ry[k];
        solve_llt(bx.data());
        for (int i = 0; i < n; i++) lx[i] = bx[i];
        for (int k = 0; k < p; k++) { double s = 0; for (int q = S.AT.p[k]; q < S.AT.p[k + 1]; q++) s += S.AT.x[q] * lx[S.AT.i[q]]; ly[k] = dinv * s; ly[k] -= dinv * ry[k]; }
        for (int k = 0; k < m; k++) { double s = 0; for (int q = S.GT.p[k]; q < S.GT.p[k + 1]; q++) s += S.GT.x[q] * lx[S.GT.i[q]]; lz[k] = s; lz[k] -= rz[k]; lz[k] *= z_reg_inv[k]; }
    }
    void eval_P_x(double alpha, const double* x, double* z) override {
        for (int i = 0; i < S.n; i++) z[i] = 0;
        for (int j = 0; j < S.n; j++) for (int q = S.P.p[j]; q < S.P.p[j + 1]; q++) { const int i = S.P.i[q]; z[i] += alpha * S.P.x[q] * x[j]; if (i != j) z[j] += alpha * S.P.x[q] * x[i]; }
    }
    void eval_A(double an, double at, const double* xn, const double* xt, double* zn, double* zt) override { csc_mv_nt(S.AT, an, at, xn, xt, zn, zt); }
    void eval_G(double an, double at, const double* xn, const double* xt, double* zn, double* zt) override { csc_mv_nt(S.GT, an, at, xn, xt, zn, zt); }
    // flop count of one factor call with the reference's own cost model (:397-418), SURVEY 8(d)
    double factor_flops() const {
        const size_t N = bi.size(); const double w = bi.back().diag; double f = 0;
        for (size_t i = 0; i + 1 < N; i++) {
            const double d = bi[i].diag, o = bi[i].off, po = i > 0 ? bi[i - 1].off : 0, pd = i > 0 ? bi[i - 1].diag : 0;
            f += po * po * pd + d * d * d / 3 + d * d * o + 2 * w * po * pd + d * d * w + w * w * d;
        }
        return f + w * w * w / 3;
    }
};

inline std::unique_ptr<KKTBackend> make_multistage_backend(const SparseMatrices& S) { return std::make_unique<MultistageKKT>(S); }

}  // namespace oracle
