// piqp_b200/csrc/dense_backend.hpp -- batched dense problem data + dense KKT backend (host-side classes).
#pragma once
#include "kkt_backend.hpp"

namespace b200 {

// Device twin of dense::Data's matrix members (include/piqp/dense/data.hpp:23-32) for a batch.
//   Pf : [batch][ld * n]  FULL symmetric P (both triangles; the reference keeps only P_utri -- the lower
//        mirror lets the KKT assembly epilogue and the row-norm sweeps read coalesced columns)
//   AT : [batch][ld * p]  (n x p column-major == A row-major)
//   GT : [batch][ld * m]
// ld = round_up(n, 8); padding rows are kept at zero.
struct DenseData {
    int batch = 0, n = 0, p = 0, m = 0, ld = 0;
    DevBuf<double> Pf, AT, GT;
    long long sP() const { return (long long)ld * n; }
    long long sA() const { return (long long)ld * p; }
    long long sG() const { return (long long)ld * m; }
    void alloc(int batch_, int n_, int p_, int m_, cudaStream_t zero_stream = 0);      // the zero-fill runs on zero_stream (0: legacy default stream, callers synchronise)
};

// element (i, j) of source instance b is src[b*sb + i*rs + j*cs]
void dense_pack_sym_upper(const double* src, long long sb, long long rs, long long cs, DenseData& D, cudaStream_t st);  // -> Pf
void dense_pack_cols(const double* src, long long sb, int rows, int cols, int ld_src, double* dst, long long sdst, int ld, int batch, cudaStream_t st);
void dense_zero_G_rows(DenseData& D, const int* row_mask /*[batch][m], 1 = zero it*/, cudaStream_t st);

// Ruiz equilibration state for a batch (dense/preconditioner.hpp:26-437)
struct RuizState {
    int batch = 0, n = 0, p = 0, m = 0;
    DevBuf<double> delta, delta_b, delta_inv, delta_b_inv, c, c_inv;   // [batch][n+p+m], [batch][n], ..., [batch]
    DevBuf<double> it, itb;                                            // per-sweep scalings
    DevBuf<int> done;                                                  // [batch] converged flag
    void alloc(int batch_, int n_, int p_, int m_);
};
// generic pieces of the equilibration shared by the dense and the sparse data layer
void fill_async(double* p, size_t n, double v, cudaStream_t st);
void ruiz_reset(RuizState& R, cudaStream_t st);
void ruiz_launch_begin(RuizState& R, cudaStream_t st);                                   // convergence test of the sweep
void ruiz_launch_finalize(RuizState& R, double* c, double* xbs, cudaStream_t st);        // limit, 1/sqrt, fold into delta / c / x_b_scaling
void ruiz_launch_inverse(RuizState& R, cudaStream_t st);
void ruiz_launch_vectors(RuizState& R, double* c, double* b, double* h_l, double* h_u, double* x_l, double* x_u, double* xbs,
                         const double* d, const double* db, const double* cs, int scale_c_and_xbs, cudaStream_t st);
void dense_ruiz_scale(DenseData& D, RuizState& R, double* c, double* b, double* h_l, double* h_u, double* x_l, double* x_u,
                      double* x_b_scaling, bool reuse_prev, bool scale_cost, int max_iter, cudaStream_t st);
void dense_ruiz_unscale(DenseData& D, RuizState& R, double* c, double* b, double* h_l, double* h_u, double* x_l, double* x_u,
                        double* x_b_scaling, cudaStream_t st);

class DenseBatchedKKT : public BatchedKKT {
public:
    DenseBatchedKKT(DenseData* data, cudaStream_t st);
    void update_data(int options) override;
    void factor(const double* delta, const double* x_reg, const double* z_reg, const int* active, int* ok) override;
    void solve(const double* rx, const double* ry, const double* rz, double* lx, double* ly, double* lz, const int* active) override;
    void eval_P_x(double alpha, const double* x, double* z, const int* active) override;
    void eval_A(double an, double at, const double* xn, const double* xt, double* zn, double* zt, const int* active) override;
    void eval_G(double an, double at, const double* xn, const double* xt, double* zn, double* zt, const int* active) override;
    void extract_P_diag(double* P_diag) override;
    void print_info() const override;
    bool graph_capturable() const override { return true; }
    double factor_flops() const override;
    double factor_bytes() const override;
    double solve_flops() const override;
    double solve_bytes() const override;

    // pieces of factor(), exposed for tests / diagnostics
    void assemble(const double* x_reg, const int* active);     // dense::KKT::update_kkt
    void cholesky(const int* active);                          // Eigen::LLT::compute
    void copy_from(const DenseBatchedKKT& o);                  // clone()
    DenseData* D;
    DevBuf<double> K;        // [batch][ld*n] assembled KKT (lower), overwritten by its Cholesky factor
    DevBuf<double> AtA;      // [batch][ld*n] lower, only if p > 0
    DevBuf<double> zinv;     // [batch][m]
    DevBuf<double> delta;    // [batch]
    DevBuf<int> fail;        // [batch] 0 = ok, else failing column + 1
    DevBuf<double> Linv;     // [batch][n/32 (padded to whole tiles)][32 x 36] inverses of the 32 x 32 diagonal blocks of L
    long long Linv_stride = 0;
    // Ozaki / tcgen05 assembly path (dense_ozaki.cuh)
    cudaStream_t chol_aux = nullptr;           // diag-tile factorisations run here, beside the block-column update on `stream`
    std::vector<cudaEvent_t> chol_ev;
    cudaStream_t chol_aux2 = nullptr;          // early part of the next diagonal tile's update (look-ahead), beside the current block column
    std::vector<cudaEvent_t> chol_ev2;
    bool chol_lookahead = true;
    ~DenseBatchedKKT() override {
        for (auto e : chol_ev) cudaEventDestroy(e);
        for (auto e : chol_ev2) cudaEventDestroy(e);
        if (chol_aux) cudaStreamDestroy(chol_aux);
        if (chol_aux2) cudaStreamDestroy(chol_aux2);
    }
    bool chol_solve64 = true;  // panel solve on 64-row half tiles with L11 read from global memory (three CTAs per SM)
    bool chol_split = true;  // Cholesky: block-column update on the two-CTA-per-SM tile kernel + solve-only panel kernel (B200_CHOL_SPLIT=0: fused panel kernel)
    bool gemm_t64 = true;    // assembly with gemm_nt_t64_kernel (two CTAs per SM) instead of gemm_nt_tile_kernel
    bool ozaki = false;
    int oz_mp = 0, oz_ntiles = 0, oz_kb = 64;   // k-step bytes: 64 (2 stages, SWIZZLE_64B) or 32 (4 stages, SWIZZLE_32B)
    DevBuf<signed char> oz_digits;   // [batch][8][n][mp]
    DevBuf<int> oz_ex, oz_tiles;     // [batch][n] row exponents; [ntiles][2] lower tiles (128-row, 64-col)
    DevBuf<double> oz_sw, oz_sc;     // [batch][m] sqrt(z_reg^-1); [batch][n] per-row output scales
    unsigned char oz_mapA[128] __attribute__((aligned(64))), oz_mapB[128] __attribute__((aligned(64)));   // CUtensorMap x 2
private:
    void compute_AtA();
    void assemble_ozaki(const double* x_reg, const int* active);
};

// C (lower 128 x 128 tiles of an n x n matrix, ld = ldc) -= A diag(w) A^T for every instance of a batch, with the DMMA tile
// kernel of the dense backend (gemm_nt_tile_kernel<EPI_SUB, true>).  A: n x K column-major (lda), w: K scale factors.
// A, C and their leading dimensions must keep 16-byte alignment of row pairs (even row offsets, even ld); rows up to
// n + (n & 1) of A must exist in memory.  Only the lower tiles of the 128-column tile columns [tj_start, tj_end) are
// updated (tj_end < 0: to the last one) and, if ncol > 0, only columns < ncol.  Used by the sparse backend's whole-GPU blocked
// LDL^T of large fronts (window / far / look-ahead updates).
void dense_syrk_sub_scaled(const double* A, long long strideA, int lda, const double* w, long long stridew, double* C, long long strideC, int ldc,
                           int n, int K, int batch, const int* active, cudaStream_t st, int tj_start = 0, int tj_end = -1, int ncol = 0);

}  // namespace b200
