// piqp_b200/csrc/sparse_data.hpp -- batched sparse problem data with a SHARED sparsity pattern (device twin of
// sparse::Data, include/piqp/sparse/data.hpp:26-54), deterministic SpMV kernels, sparse Ruiz sweeps.
//
// All instances of a batch share the patterns of P_utri (upper triangle, CSC), AT (n x p, CSC) and GT (n x m, CSC);
// values are instance-major [batch][nnz].  Every pattern also carries its CSR view (row pointers, column indices and
// the position of each entry in the CSC value array), so both M*x and M^T*x are gather-type sums: one thread per
// output element, fixed summation order, no atomics.
#pragma once
#include <vector>
#include "dense_backend.hpp"   // RuizState + shared Ruiz helpers

namespace b200 {

struct Pattern {
    int rows = 0, cols = 0, nnz = 0;
    std::vector<int> p, i;                 // host CSC
    std::vector<int> rp, ci, pos, colof;   // host CSR view (+ position into the CSC value array), column of each CSC entry
    DevBuf<int> d_p, d_i, d_rp, d_ci, d_pos, d_colof;
    void build(int rows_, int cols_, const int* cp, const int* ri);   // copies, builds CSR view, uploads
};

struct SparseData {
    int batch = 0, n = 0, p = 0, m = 0;
    Pattern P, AT, GT;
    DevBuf<double> Px, ATx, GTx;           // [batch][nnz]
    void alloc_values(int batch_);
};

// SpMV (deterministic gathers).  out[b][*] (+)= alpha * ...
void spmv_rows(const Pattern& M, const double* vals, double alpha, const double* x, int x_len, double* out, int accumulate,
               const double* col_scale, const double* alpha_v, int alpha_v_inverse, int batch, const int* active, cudaStream_t st);   // out[i] over rows of M, x indexed by column
void spmv_cols(const Pattern& M, const double* vals, double alpha, const double* x, int x_len, double* out,
               const double* sub, double alpha2, const double* out_scale, const double* alpha_v, int alpha_v_inverse,
               int batch, const int* active, cudaStream_t st);                                                                // out[k] over columns of M, x indexed by row
void spmv_sym_upper(const Pattern& P, const double* vals, double alpha, const double* x, double* out, int batch, const int* active, cudaStream_t st);
void sparse_extract_diag(const SparseData& S, double* P_diag, cudaStream_t st);
void sparse_zero_G_rows(SparseData& S, const int* row_mask, cudaStream_t st);

void sparse_ruiz_scale(SparseData& S, RuizState& R, double* c, double* b, double* h_l, double* h_u, double* x_l, double* x_u,
                       double* xbs, bool reuse_prev, bool scale_cost, int max_iter, cudaStream_t st);
void sparse_ruiz_unscale(SparseData& S, RuizState& R, double* c, double* b, double* h_l, double* h_u, double* x_l, double* x_u,
                         double* xbs, cudaStream_t st);

}  // namespace b200
