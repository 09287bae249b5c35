// piqp_b200/csrc/sparse_data.cu -- see sparse_data.hpp
#include "sparse_data.hpp"
#include <algorithm>

namespace b200 {

void Pattern::build(int rows_, int cols_, const int* cp, const int* ri) {
    rows = rows_; cols = cols_;
    p.assign(cols + 1, 0);
    if (cp) p.assign(cp, cp + cols + 1);
    nnz = p[cols];
    i.assign(ri, ri + nnz);
    colof.assign(nnz, 0);
    rp.assign(rows + 1, 0);
    for (int j = 0; j < cols; j++) for (int q = p[j]; q < p[j + 1]; q++) { colof[q] = j; rp[i[q] + 1]++; }
    for (int r = 0; r < rows; r++) rp[r + 1] += rp[r];
    ci.assign(nnz, 0); pos.assign(nnz, 0);
    std::vector<int> w(rp.begin(), rp.end() - 1);
    for (int j = 0; j < cols; j++) for (int q = p[j]; q < p[j + 1]; q++) { const int t = w[i[q]]++; ci[t] = j; pos[t] = q; }
    auto up = [](DevBuf<int>& d, const std::vector<int>& h) { d.alloc(std::max<size_t>(h.size(), 1)); if (!h.empty()) B200_CUDA(cudaMemcpy(d.get(), h.data(), h.size() * sizeof(int), cudaMemcpyHostToDevice)); };
    up(d_p, p); up(d_i, i); up(d_rp, rp); up(d_ci, ci); up(d_pos, pos); up(d_colof, colof);
}

void SparseData::alloc_values(int batch_) {
    batch = batch_;
    Px.alloc(std::max<size_t>((size_t)batch * P.nnz, 1)); ATx.alloc(std::max<size_t>((size_t)batch * AT.nnz, 1)); GTx.alloc(std::max<size_t>((size_t)batch * GT.nnz, 1));
}

// ---------------------------------------------------------------------------------------------------
__global__ void spmv_rows_kernel(const int* rp, const int* ci, const int* pos, int rows, int nnz, const double* vals, double alpha,
                                 const double* x, int x_len, double* out, int accumulate, const double* col_scale,
                                 const double* alpha_v, int alpha_v_inverse, const int* active) {
    const int b = blockIdx.y;
    if (active && !active[b]) return;
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const double* v = vals + (size_t)b * nnz;
    const double* xb = x + (size_t)b * x_len;
    const double* cs = col_scale ? col_scale + (size_t)b * x_len : nullptr;
    double acc = 0.0;
    // unrolled by 4: the index -> value loads of four entries are in flight together (one L2 round trip per group instead of per entry);
    // the products are still accumulated one after the other in row order, so the sum is bit-identical to the rolled loop
    const int q0 = rp[r], q1 = rp[r + 1];
    if (cs) { for (int q = q0; q < q1; q++) { const int c = ci[q]; acc += v[pos[q]] * (xb[c] * cs[c]); } }
    else {
        int q = q0;
        for (; q + 4 <= q1; q += 4) {
            const int c0 = ci[q], c1 = ci[q + 1], c2 = ci[q + 2], c3 = ci[q + 3], p0 = pos[q], p1 = pos[q + 1], p2 = pos[q + 2], p3 = pos[q + 3];
            const double x0 = xb[c0], x1 = xb[c1], x2 = xb[c2], x3 = xb[c3], v0 = v[p0], v1 = v[p1], v2 = v[p2], v3 = v[p3];
            acc += v0 * x0; acc += v1 * x1; acc += v2 * x2; acc += v3 * x3;
        }
        for (; q < q1; q++) acc += v[pos[q]] * xb[ci[q]];
    }
    double al = alpha;
    if (alpha_v) al *= alpha_v_inverse ? 1.0 / alpha_v[b] : alpha_v[b];
    double* o = out + (size_t)b * rows;
    if (accumulate) o[r] += al * acc; else o[r] = al * acc;
}
void spmv_rows(const Pattern& M, const double* vals, double alpha, const double* x, int x_len, double* out, int accumulate,
               const double* col_scale, const double* alpha_v, int alpha_v_inverse, int batch, const int* active, cudaStream_t st) {
    if (M.rows == 0) return;
    dim3 g(ceil_div(M.rows, 128), batch);
    B200_LAUNCH(spmv_rows_kernel, g, 128, 0, st, M.d_rp.get(), M.d_ci.get(), M.d_pos.get(), M.rows, M.nnz, vals, alpha, x, x_len, out, accumulate,
                col_scale, alpha_v, alpha_v_inverse, active);
}

__global__ void spmv_cols_kernel(const int* cp, const int* ri, int rows, int cols, int nnz, const double* vals, double alpha, const double* x,
                                 double* out, const double* sub, double alpha2, const double* out_scale, const double* alpha_v,
                                 int alpha_v_inverse, const int* active) {
    const int b = blockIdx.y;
    if (active && !active[b]) return;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= cols) return;
    const double* v = vals + (size_t)b * nnz;
    const double* xb = x + (size_t)b * rows;
    double acc = 0.0;
    {
        const int q0 = cp[k], q1 = cp[k + 1];
        int q = q0;
        for (; q + 4 <= q1; q += 4) {      // see spmv_rows_kernel
            const int r0 = ri[q], r1 = ri[q + 1], r2 = ri[q + 2], r3 = ri[q + 3];
            const double v0 = v[q], v1 = v[q + 1], v2 = v[q + 2], v3 = v[q + 3], x0 = xb[r0], x1 = xb[r1], x2 = xb[r2], x3 = xb[r3];
            acc += v0 * x0; acc += v1 * x1; acc += v2 * x2; acc += v3 * x3;
        }
        for (; q < q1; q++) acc += v[q] * xb[ri[q]];
    }
    double al = alpha, al2 = alpha2;
    if (alpha_v) { const double s = alpha_v_inverse ? 1.0 / alpha_v[b] : alpha_v[b]; al *= s; al2 *= s; }
    double r = al * acc;
    if (sub) r -= al2 * sub[(size_t)b * cols + k];
    if (out_scale) r *= out_scale[(size_t)b * cols + k];
    out[(size_t)b * cols + k] = r;
}
void spmv_cols(const Pattern& M, const double* vals, double alpha, const double* x, int, double* out, const double* sub, double alpha2,
               const double* out_scale, const double* alpha_v, int alpha_v_inverse, int batch, const int* active, cudaStream_t st) {
    if (M.cols == 0) return;
    dim3 g(ceil_div(M.cols, 128), batch);
    B200_LAUNCH(spmv_cols_kernel, g, 128, 0, st, M.d_p.get(), M.d_i.get(), M.rows, M.cols, M.nnz, vals, alpha, x, out, sub, alpha2, out_scale,
                alpha_v, alpha_v_inverse, active);
}

// z = alpha * P x with P given by its upper triangle (sparse/kkt.hpp:179-185)
__global__ void spmv_sym_upper_kernel(const int* cp, const int* ri, const int* rp, const int* ci, const int* pos, int n, int nnz,
                                      const double* vals, double alpha, const double* x, double* out, const int* active) {
    const int b = blockIdx.y;
    if (active && !active[b]) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* v = vals + (size_t)b * nnz;
    const double* xb = x + (size_t)b * n;
    double acc = 0.0;
    for (int q = rp[i]; q < rp[i + 1]; q++) acc += v[pos[q]] * xb[ci[q]];                          // row i of the upper triangle (j >= i)
    for (int q = cp[i]; q < cp[i + 1]; q++) { const int r = ri[q]; if (r < i) acc += v[q] * xb[r]; }   // column i above the diagonal
    out[(size_t)b * n + i] = alpha * acc;
}
void spmv_sym_upper(const Pattern& P, const double* vals, double alpha, const double* x, double* out, int batch, const int* active, cudaStream_t st) {
    if (P.rows == 0) return;
    dim3 g(ceil_div(P.rows, 128), batch);
    B200_LAUNCH(spmv_sym_upper_kernel, g, 128, 0, st, P.d_p.get(), P.d_i.get(), P.d_rp.get(), P.d_ci.get(), P.d_pos.get(), P.rows, P.nnz, vals, alpha, x, out, active);
}

__global__ void sparse_diag_kernel(const int* cp, const int* ri, int n, int nnz, const double* vals, double* out) {
    const int b = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    double d = 0.0;
    for (int q = cp[j]; q < cp[j + 1]; q++) if (ri[q] == j) d = vals[(size_t)b * nnz + q];
    out[(size_t)b * n + j] = d;
}
void sparse_extract_diag(const SparseData& S, double* P_diag, cudaStream_t st) {
    if (S.n == 0) return;
    dim3 g(ceil_div(S.n, 128), S.batch);
    B200_LAUNCH(sparse_diag_kernel, g, 128, 0, st, S.P.d_p.get(), S.P.d_i.get(), S.n, S.P.nnz, S.Px.get(), P_diag);
}

__global__ void sparse_zero_rows_kernel(const int* colof, int nnz, int m, double* vals, const int* mask) {
    const int b = blockIdx.y;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < nnz && mask[(size_t)b * m + colof[q]]) vals[(size_t)b * nnz + q] = 0.0;
}
void sparse_zero_G_rows(SparseData& S, const int* row_mask, cudaStream_t st) {   // sparse/data.hpp:212-216
    if (S.GT.nnz == 0) return;
    dim3 g(ceil_div(S.GT.nnz, 256), S.batch);
    B200_LAUNCH(sparse_zero_rows_kernel, g, 256, 0, st, S.GT.d_colof.get(), S.GT.nnz, S.m, S.GTx.get(), row_mask);
}

// ---------------------------------------------------------------------------------------------------
// sparse Ruiz (sparse/preconditioner.hpp:64-290), same arithmetic order as the reference / oracle
// ---------------------------------------------------------------------------------------------------
__global__ void sruiz_varnorm_kernel(const int* Pcp, const int* Pri, const int* Prp, const int* Ppos, int Pnnz, const double* Px,
                                     const int* Arp, const int* Apos, int Annz, const double* Ax,
                                     const int* Grp, const int* Gpos, int Gnnz, const double* Gx,
                                     int n, int N, const double* xbs, double* it, double* itb, const int* done) {
    const int b = blockIdx.y;
    if (done[b]) return;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    double v = 0.0;
    const double* px = Px + (size_t)b * Pnnz;
    for (int q = Pcp[j]; q < Pcp[j + 1]; q++) v = fmax(v, fabs(px[q]));                 // column j of the upper triangle
    for (int q = Prp[j]; q < Prp[j + 1]; q++) v = fmax(v, fabs(px[Ppos[q]]));           // row j of the upper triangle
    const double* ax = Ax + (size_t)b * Annz;
    for (int q = Arp[j]; q < Arp[j + 1]; q++) v = fmax(v, fabs(ax[Apos[q]]));
    const double* gx = Gx + (size_t)b * Gnnz;
    for (int q = Grp[j]; q < Grp[j + 1]; q++) v = fmax(v, fabs(gx[Gpos[q]]));
    const double xb = xbs[(size_t)b * n + j];
    it[(size_t)b * N + j] = fmax(v, xb);
    itb[(size_t)b * n + j] = xb;
}
__global__ void sruiz_connorm_kernel(const int* Acp, int Annz, const double* Ax, const int* Gcp, int Gnnz, const double* Gx,
                                     int n, int p, int m, int N, double* it, const int* done) {
    const int b = blockIdx.y;
    if (done[b]) return;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= p + m) return;
    double v = 0.0;
    if (k < p) { const double* ax = Ax + (size_t)b * Annz; for (int q = Acp[k]; q < Acp[k + 1]; q++) v = fmax(v, fabs(ax[q])); }
    else { const int kk = k - p; const double* gx = Gx + (size_t)b * Gnnz; for (int q = Gcp[kk]; q < Gcp[kk + 1]; q++) v = fmax(v, fabs(gx[q])); }
    it[(size_t)b * N + n + k] = v;
}
// vals <- gamma * vals, then pre-multiply by the row scaling and post-multiply by the column scaling (utils.hpp:171-201)
__global__ void sruiz_scale_kernel(const int* ri, const int* colof, int nnz, double* vals, const double* d, int N, int row_off, int col_off,
                                   const double* gamma, const int* done) {
    const int b = blockIdx.y;
    if (done && done[b]) return;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nnz) return;
    const double* dv = d + (size_t)b * N;
    double v = vals[(size_t)b * nnz + q];
    if (gamma) v *= gamma[b];
    v *= dv[row_off + ri[q]];
    v *= dv[col_off + colof[q]];
    vals[(size_t)b * nnz + q] = v;
}
// cost scaling (sparse/preconditioner.hpp:141-175), one CTA per instance (default off)
__global__ void sruiz_cost_kernel(const int* Pcp, const int* Pri, const int* Prp, const int* Ppos, int Pnnz, double* Px, int n, double* c,
                                  double* cscale, const int* done) {
    __shared__ double red[64];
    const int b = blockIdx.x;
    if (done[b]) return;
    double* px = Px + (size_t)b * Pnnz;
    double v[2] = {0.0, 0.0};
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        double r = 0.0;
        for (int q = Pcp[j]; q < Pcp[j + 1]; q++) r = fmax(r, fabs(px[q]));
        for (int q = Prp[j]; q < Prp[j + 1]; q++) r = fmax(r, fabs(px[Ppos[q]]));
        v[0] += r; v[1] = fmax(v[1], fabs(c[(size_t)b * n + j]));
    }
    const int op[2] = {RED_SUM, RED_MAX};
    block_reduce<2>(v, op, red);
    double g = v[0] / double(n);
    g = g < 1e-4 ? 1.0 : (g > 1e4 ? 1e4 : g);
    g = fmax(g, v[1]);
    g = g < 1e-4 ? 1.0 : (g > 1e4 ? 1e4 : g);
    g = 1.0 / g;
    for (int q = threadIdx.x; q < Pnnz; q += blockDim.x) px[q] *= g;
    for (int j = threadIdx.x; j < n; j += blockDim.x) c[(size_t)b * n + j] *= g;
    if (threadIdx.x == 0) cscale[b] *= g;
}

static void scale_all(SparseData& S, const double* d, const double* gamma, const int* done, cudaStream_t st) {
    const int N = S.n + S.p + S.m, B = S.batch;
    if (S.P.nnz) { dim3 g(ceil_div(S.P.nnz, 256), B);
        B200_LAUNCH(sruiz_scale_kernel, g, 256, 0, st, S.P.d_i.get(), S.P.d_colof.get(), S.P.nnz, S.Px.get(), d, N, 0, 0, gamma, done); }
    if (S.AT.nnz) { dim3 g(ceil_div(S.AT.nnz, 256), B);
        B200_LAUNCH(sruiz_scale_kernel, g, 256, 0, st, S.AT.d_i.get(), S.AT.d_colof.get(), S.AT.nnz, S.ATx.get(), d, N, 0, S.n, (const double*)nullptr, done); }
    if (S.GT.nnz) { dim3 g(ceil_div(S.GT.nnz, 256), B);
        B200_LAUNCH(sruiz_scale_kernel, g, 256, 0, st, S.GT.d_i.get(), S.GT.d_colof.get(), S.GT.nnz, S.GTx.get(), d, N, 0, S.n + S.p, (const double*)nullptr, done); }
}

void sparse_ruiz_scale(SparseData& S, RuizState& R, double* c, double* b, double* h_l, double* h_u, double* x_l, double* x_u,
                       double* xbs, bool reuse_prev, bool scale_cost, int max_iter, cudaStream_t st) {
    const int n = S.n, p = S.p, m = S.m, N = n + p + m, B = S.batch;
    if (!reuse_prev) {
        ruiz_reset(R, st);
        for (int iter = 0; iter < max_iter; iter++) {
            ruiz_launch_begin(R, st);
            if (n > 0) { dim3 g(ceil_div(n, 128), B);
                B200_LAUNCH(sruiz_varnorm_kernel, g, 128, 0, st, S.P.d_p.get(), S.P.d_i.get(), S.P.d_rp.get(), S.P.d_pos.get(), S.P.nnz, S.Px.get(),
                            S.AT.d_rp.get(), S.AT.d_pos.get(), S.AT.nnz, S.ATx.get(), S.GT.d_rp.get(), S.GT.d_pos.get(), S.GT.nnz, S.GTx.get(),
                            n, N, xbs, R.it.get(), R.itb.get(), R.done.get()); }
            if (p + m > 0) { dim3 g(ceil_div(p + m, 128), B);
                B200_LAUNCH(sruiz_connorm_kernel, g, 128, 0, st, S.AT.d_p.get(), S.AT.nnz, S.ATx.get(), S.GT.d_p.get(), S.GT.nnz, S.GTx.get(), n, p, m, N,
                            R.it.get(), R.done.get()); }
            ruiz_launch_finalize(R, c, xbs, st);
            scale_all(S, R.it.get(), nullptr, R.done.get(), st);
            if (scale_cost)
                B200_LAUNCH(sruiz_cost_kernel, B, 256, 0, st, S.P.d_p.get(), S.P.d_i.get(), S.P.d_rp.get(), S.P.d_pos.get(), S.P.nnz, S.Px.get(), n, c, R.c.get(), R.done.get());
        }
        ruiz_launch_inverse(R, st);
        ruiz_launch_vectors(R, c, b, h_l, h_u, x_l, x_u, xbs, R.delta.get(), R.delta_b.get(), R.c.get(), 0, st);
    } else {
        scale_all(S, R.delta.get(), R.c.get(), nullptr, st);
        ruiz_launch_vectors(R, c, b, h_l, h_u, x_l, x_u, xbs, R.delta.get(), R.delta_b.get(), R.c.get(), 1, st);
    }
}
void sparse_ruiz_unscale(SparseData& S, RuizState& R, double* c, double* b, double* h_l, double* h_u, double* x_l, double* x_u,
                         double* xbs, cudaStream_t st) {
    scale_all(S, R.delta_inv.get(), R.c_inv.get(), nullptr, st);
    ruiz_launch_vectors(R, c, b, h_l, h_u, x_l, x_u, xbs, R.delta_inv.get(), R.delta_b_inv.get(), R.c_inv.get(), 1, st);
}

}  // namespace b200
