// piqp_b200/csrc/multistage_backend.cu -- see multistage_backend.hpp
#include "multistage_backend.hpp"
#include "multistage_chain.cuh"
#include "multistage_partition.cuh"
#include <cmath>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>

namespace b200 {

// =====================================================================================================
// host: structure detection (multistage_kkt.hpp:420-597) and index maps
// =====================================================================================================
namespace {
typedef unsigned long long u64;
inline u64 fl_gemm(u64 m, u64 n, u64 k) { return 2 * m * n * k; }   // :397-418
inline u64 fl_trsm(u64 m, u64 n) { return m * m * n; }
inline u64 fl_syrk(u64 n, u64 k) { return n * n * k; }
inline u64 fl_potrf(u64 n) { return n * n * n / 3; }
}  // namespace

bool MsStructure::detect(const Pattern& P, const Pattern& AT, const Pattern& GT) {
    B200_ZONE("piqp::MultistageKKT::extract_arrow_structure");
    n = P.rows;
    // structural upper pattern of C = P + I + A^T A + G^T G, row by row (sorted, unique)
    std::vector<std::vector<int>> up(n);
    for (int j = 0; j < n; j++) for (int q = P.p[j]; q < P.p[j + 1]; q++) if (P.i[q] <= j) up[P.i[q]].push_back(j);
    for (int i = 0; i < n; i++) up[i].push_back(i);
    for (const Pattern* M : {&AT, &GT})
        for (int r = 0; r < M->cols; r++)
            for (int a = M->p[r]; a < M->p[r + 1]; a++)
                for (int b = a; b < M->p[r + 1]; b++) { const int lo = std::min(M->i[a], M->i[b]), hi = std::max(M->i[a], M->i[b]); up[lo].push_back(hi); }
    for (auto& v : up) { std::sort(v.begin(), v.end()); v.erase(std::unique(v.begin(), v.end()), v.end()); }

    struct St { int prev_diag = 0, start = 0, diag = 0, off = 0, arrow = 0; } cur;
    u64 f_tri = 0, f_a_nosyrk = 0, f_a_syrk = 0;
    auto advance = [&](int row, St s) {
        if (row >= n) return s;
        for (int col : up[row]) {
            if (!(col >= s.start && col + s.arrow < n)) continue;
            const int size_now = s.diag + s.off;
            const int size_new = std::max(col - s.start + 1, size_now);
            const int diag_cap = row - s.start + 1;
            const int diag_new = std::max(std::max(s.diag, (size_new + 1) / 2), diag_cap);
            const int off_new = size_new - diag_new;
            const int arrow_new = std::min(std::max(s.arrow, n - col), n - s.start - s.diag - s.off);
            const u64 tri_new = f_tri + fl_syrk((u64)diag_new, (u64)s.prev_diag) + fl_potrf((u64)diag_new) + fl_trsm((u64)diag_new, (u64)off_new);
            const u64 aw = (u64)((s.arrow + 3) / 4) * 4, awn = (u64)((arrow_new + 3) / 4) * 4;   // kernels work on multiples of 4
            const u64 arrow_now = aw * f_a_nosyrk + aw * aw * f_a_syrk + fl_potrf(aw);
            const u64 arrow_then = awn * f_a_nosyrk + awn * awn * f_a_syrk + fl_gemm(awn, (u64)s.prev_diag, (u64)diag_new) +
                                   fl_trsm((u64)diag_new, awn) + fl_syrk(awn, (u64)diag_new) + fl_potrf(awn);
            if (tri_new - f_tri <= arrow_then - arrow_now) { s.diag = diag_new; s.off = off_new; }
            else s.arrow = arrow_new;
        }
        return s;
    };
    bi.clear();
    for (int i = 0; i < n; i++) {
        cur = advance(i, cur);
        if (i + 1 < cur.start + cur.diag) continue;
        const bool ratio_ok = cur.diag >= 2 * cur.off;
        const bool at_end = i + 1 >= n - cur.arrow;
        bool grows = false;
        if (!ratio_ok && !at_end) { St nx = advance(i + 1, cur); grows = nx.diag + nx.off > cur.diag + cur.off; }
        if (ratio_ok || at_end || grows) {
            bi.push_back({cur.start, cur.diag, cur.off});
            f_tri += fl_syrk((u64)cur.diag, (u64)(cur.prev_diag + 1)) + fl_potrf((u64)cur.diag) + fl_trsm((u64)cur.diag, (u64)cur.off);
            f_a_nosyrk += fl_gemm(1, (u64)cur.prev_diag, (u64)cur.diag) + fl_trsm((u64)cur.diag, 1);
            f_a_syrk += fl_syrk(1, (u64)cur.diag);
            cur.start += cur.diag; cur.prev_diag = cur.diag; cur.diag = cur.off; cur.off = 0;
        }
        if (at_end && cur.diag > 0) {
            bi.push_back({cur.start, cur.diag, cur.off});
            cur.start += cur.diag; cur.prev_diag = cur.diag; cur.diag = cur.off; cur.off = 0;
        }
        if (at_end) break;
    }
    for (size_t i = 0; i + 1 < bi.size(); i++)   // merge blocks that were split in two (:569-579)
        if (bi[i].off == bi[i + 1].diag && bi[i + 1].off == 0) { bi[i].diag += bi[i].off; bi[i].off = 0; bi.erase(bi.begin() + (long)i + 1); }
    bi.push_back({cur.start, cur.arrow, 0});
    N = (int)bi.size(); w = bi.back().diag;
    if (N < 2) { error = "multistage: degenerate block structure"; return false; }

    // storage layout
    offD.assign(N, 0); offB.assign(N, 0); offE.assign(N, 0); offI.assign(N, 0);
    total = 0; total_inv = 0; dmax = 1; omax = 1;
    for (int i = 0; i < N; i++) {
        const int d = bi[i].diag, o = (i + 2 < N) ? bi[i].off : 0;
        offD[i] = total; total += d * d;
        offB[i] = total; total += o * d;
        offE[i] = total; total += (i + 1 < N) ? w * d : 0;
        offI[i] = total_inv; total_inv += d * d;
        dmax = std::max(dmax, d); omax = std::max(omax, o);
    }
    blk_of.assign(n, 0);
    for (int b = 0; b < N; b++) for (int k = 0; k < bi[b].diag; k++) blk_of[bi[b].start + k] = b;

    // maps
    P_slot.assign(P.nnz, -1);
    for (int j = 0; j < n; j++) for (int q = P.p[j]; q < P.p[j + 1]; q++) {
        if (P.i[q] > j) continue;
        const int s = slot(j, P.i[q]);
        if (s < 0) { error = "multistage: an entry of P lies outside the detected block structure"; return false; }
        P_slot[q] = s;
    }
    diag_slot.assign(n, 0);
    for (int i = 0; i < n; i++) diag_slot[i] = slot(i, i);
    auto build = [&](const Pattern& M, std::vector<int>& ptr, std::vector<int>& qa, std::vector<int>& qb, std::vector<int>* row) {
        std::vector<int> cnt(total + 1, 0);
        for (int r = 0; r < M.cols; r++) for (int a = M.p[r]; a < M.p[r + 1]; a++) for (int b = M.p[r]; b < M.p[r + 1]; b++) {
            if (M.i[a] < M.i[b]) continue;
            const int s = slot(M.i[a], M.i[b]);
            if (s < 0) return false;
            cnt[s + 1]++;
        }
        ptr.assign(total + 1, 0);
        for (int s = 0; s < total; s++) ptr[s + 1] = ptr[s] + cnt[s + 1];
        qa.assign(ptr[total], 0); qb.assign(ptr[total], 0);
        if (row) row->assign(ptr[total], 0);
        std::vector<int> wpos(ptr.begin(), ptr.end() - 1);
        for (int r = 0; r < M.cols; r++) for (int a = M.p[r]; a < M.p[r + 1]; a++) for (int b = M.p[r]; b < M.p[r + 1]; b++) {
            if (M.i[a] < M.i[b]) continue;
            const int s = slot(M.i[a], M.i[b]);
            const int t = wpos[s]++;
            qa[t] = a; qb[t] = b;
            if (row) (*row)[t] = r;
        }
        return true;
    };
    if (!build(AT, a_ptr, a_qa, a_qb, nullptr)) { error = "multistage: a row of A couples variables outside the detected block structure"; return false; }
    if (!build(GT, g_ptr, g_qa, g_qb, &g_row)) { error = "multistage: a row of G couples variables outside the detected block structure"; return false; }
    return true;
}

int MsStructure::slot(int i, int j) const {
    const int bj = blk_of[j], bI = blk_of[i];
    if (bI == bj) return offD[bj] + (i - bi[bj].start) + (j - bi[bj].start) * bi[bj].diag;
    if (w > 0 && i >= n - w) return offE[bj] + (i - (n - w)) + (j - bi[bj].start) * w;
    if (bI == bj + 1 && bj + 2 < N && i - bi[bI].start < bi[bj].off) return offB[bj] + (i - bi[bI].start) + (j - bi[bj].start) * bi[bj].off;
    return -1;
}
double MsStructure::factor_flops() const {   // the BLAS calls factor_kkt issues, with the reference's cost model
    double f = 0;
    for (int i = 0; i + 1 < N; i++) {
        const double d = bi[i].diag, o = (i + 2 < N) ? bi[i].off : 0, po = i > 0 ? bi[i - 1].off : 0, pd = i > 0 ? bi[i - 1].diag : 0, W = w;
        f += po * po * pd + d * d * d / 3 + d * d * o + 2 * W * po * pd + d * d * W + W * W * d;
    }
    return f + (double)w * w * w / 3;
}
double MsStructure::factor_bytes() const { return 8.0 * 3.0 * total; }   // P/AtA blocks read, kkt_fac written and read once
double MsStructure::solve_flops() const {
    double f = 2.0 * w * w;
    for (int i = 0; i + 1 < N; i++) { const double d = bi[i].diag, o = (i + 2 < N) ? bi[i].off : 0; f += 2 * (d * d + 2 * o * d + 2 * w * d); }
    return f;
}
double MsStructure::solve_bytes() const {
    double by = 0;
    for (int i = 0; i + 1 < N; i++) { const double d = bi[i].diag, o = (i + 2 < N) ? bi[i].off : 0; by += 8.0 * 2 * (d * d / 2 + o * d + w * d); }
    return by + 8.0 * 5 * n;
}

// =====================================================================================================
// device kernels
// =====================================================================================================
// out[slot] = sum_list (w[row] *) vals[qa] * vals[qb]
__global__ void ms_accumulate_kernel(const int* ptr, const int* qa, const int* qb, const int* row, int total, int nnz, int m, const double* vals,
                                     const double* w, double* out, const int* active) {
    const int b = blockIdx.y;
    if (active && !active[b]) return;
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= total) return;
    const double* v = vals + (size_t)b * nnz;
    const double* wb = w ? w + (size_t)b * m : nullptr;
    double acc = 0.0;
    for (int t = ptr[s]; t < ptr[s + 1]; t++) { double a = v[qa[t]]; if (wb) a *= wb[row[t]]; acc += a * v[qb[t]]; }
    out[(size_t)b * total + s] = acc;
}
__global__ void ms_scatter_P_kernel(const int* P_slot, int nnz, int total, const double* Px, double* Pblk) {
    const int b = blockIdx.y;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < nnz && P_slot[q] >= 0) Pblk[(size_t)b * total + P_slot[q]] = Px[(size_t)b * nnz + q];
}
// kkt_fac = P + AtA / delta + GtG + diag(x_reg)   (construct_kkt_fac :1008-1219 fused with block_syrk_ln of the scaled G)
__global__ void ms_assemble_kernel(const int* gptr, const int* gqa, const int* gqb, const int* grow, const int* diag_var, int total, int gnnz, int m, int n,
                                   const double* Pblk, const double* AtAblk, const double* Gx, const double* zinv, const double* delta,
                                   const double* x_reg, double* fac, const int* active) {
    const int b = blockIdx.y;
    if (active && !active[b]) return;
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= total) return;
    const double* g = Gx + (size_t)b * gnnz;
    const double* wz = zinv + (size_t)b * m;
    double gtg = 0.0;
    for (int t = gptr[s]; t < gptr[s + 1]; t++) gtg += (g[gqa[t]] * wz[grow[t]]) * g[gqb[t]];
    double v = Pblk[(size_t)b * total + s];
    v += (1.0 / delta[b]) * AtAblk[(size_t)b * total + s];
    v += gtg;
    const int dv = diag_var[s];
    if (dv >= 0) v += x_reg[(size_t)b * n + dv];
    fac[(size_t)b * total + s] = v;
}

constexpr int MS_T = 128;

// in-smem Cholesky of the d x d block A (ld = d), one barrier per column; then inv <- L^{-1} (dense d x d, ld = d)
__device__ void ms_potrf_inv(double* A, double* inv, double* rinv, int d) {
    const int tid = threadIdx.x;
    for (int k = 0; k < d; k++) {
        __syncthreads();
        const double dk = A[k + k * d];
        const double r = rsqrt(dk);
        if (tid == 0) rinv[k] = r;
        const int rem = d - k - 1;
        for (int e = tid; e < rem * rem; e += MS_T) {      // trailing lower update with the unscaled column k
            const int c = k + 1 + e / rem, rr = k + 1 + e % rem;
            if (rr >= c) A[rr + c * d] -= (A[rr + k * d] * r) * (A[c + k * d] * r);
        }
    }
    __syncthreads();
    for (int e = tid; e < d * d; e += MS_T) {
        const int c = e / d, rr = e % d;
        if (rr > c) A[e] *= rinv[c]; else if (rr == c) A[e] = A[e] * rinv[c]; else A[e] = 0.0;    // l_cc = d_cc * rsqrt(d_cc)
    }
    __syncthreads();
    // inverse: thread c computes column c of L^{-1}
    for (int c = tid; c < d; c += MS_T) {
        for (int i = 0; i < d; i++) {
            double s = (i == c) ? 1.0 : 0.0;
            for (int k = c; k < i; k++) s -= A[i + k * d] * inv[k + c * d];
            inv[i + c * d] = (i < c) ? 0.0 : s * rinv[i];
        }
    }
    __syncthreads();
}
// X (rows x d, ld = ldx) <- X * L^{-T} = X * inv^T : X[r, j] = sum_k X[r,k] inv[j,k]; rows processed thread-per-row via a scratch row in registers-free manner
__device__ void ms_apply_invT(double* X, int rows, int ldx, const double* inv, int d, double* scratch) {
    // scratch: rows x d temporary (smem)
    for (int e = threadIdx.x; e < rows * d; e += MS_T) {
        const int r = e % rows, j = e / rows;
        double s = 0.0;
        for (int k = 0; k <= j; k++) s += X[r + k * ldx] * inv[j + k * d];
        scratch[r + j * rows] = s;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < rows * d; e += MS_T) { const int r = e % rows, j = e / rows; X[r + j * ldx] = scratch[r + j * rows]; }
    __syncthreads();
}

// factor_kkt (:1253-1352): one CTA per instance, the chain of stages processed in shared memory
__global__ void __launch_bounds__(MS_T) ms_factor_kernel(MsDev s, double* fac_all, double* inv_all, const int* active) {
    extern __shared__ __align__(16) double sm[];
    const int b = blockIdx.x;
    if (active && !active[b]) return;
    double* fac = fac_all + (size_t)b * s.total;
    double* Linv = inv_all + (size_t)b * s.total_inv;
    const int dm = s.dmax, om = s.omax, w = s.w, tid = threadIdx.x;
    double* Dc = sm;                       // dm*dm
    double* Ic = Dc + dm * dm;             // dm*dm (inverse of the current pivot block)
    double* Cb0 = Ic + dm * dm;            // om*dm x2
    double* Cb1 = Cb0 + om * dm;
    double* Fb0 = Cb1 + om * dm;           // w*dm x2
    double* Fb1 = Fb0 + w * dm;
    double* DN = Fb1 + w * dm;             // w*w
    double* scr = DN + w * w;              // max(om, w) * dm
    double* rinv = scr + (om > w ? om : w) * dm;   // max(dm, w)
    const int N = s.N;
    if (w > 0) for (int e = tid; e < w * w; e += MS_T) DN[e] = fac[s.offD[N - 1] + e];
    double *Cprev = Cb1, *Ccur = Cb0, *Fprev = Fb1, *Fcur = Fb0;
    for (int i = 0; i + 1 < N; i++) {
        const int d = s.diag[i], o = (i + 2 < N) ? s.off[i] : 0;
        const int po = i > 0 ? s.off[i - 1] : 0, pd = i > 0 ? s.diag[i - 1] : 0;
        for (int e = tid; e < d * d; e += MS_T) Dc[e] = fac[s.offD[i] + e];
        for (int e = tid; e < o * d; e += MS_T) Ccur[e] = fac[s.offB[i] + e];
        for (int e = tid; e < w * d; e += MS_T) Fcur[e] = fac[s.offE[i] + e];
        __syncthreads();
        if (po > 0) {
            // D_i(0:po,0:po) -= C_{i-1} C_{i-1}^T ; E_i(:,0:po) -= F_{i-1} C_{i-1}^T
            for (int e = tid; e < po * po; e += MS_T) {
                const int c = e / po, r = e % po;
                if (r < c) continue;
                double acc = 0.0;
                for (int k = 0; k < pd; k++) acc += Cprev[r + k * po] * Cprev[c + k * po];
                Dc[r + c * d] -= acc;
            }
            for (int e = tid; e < w * po; e += MS_T) {
                const int c = e / w, r = e % w;
                double acc = 0.0;
                for (int k = 0; k < pd; k++) acc += Fprev[r + k * w] * Cprev[c + k * po];
                Fcur[r + c * w] -= acc;
            }
        }
        ms_potrf_inv(Dc, Ic, rinv, d);
        if (o > 0) ms_apply_invT(Ccur, o, o, Ic, d, scr);
        if (w > 0) {
            ms_apply_invT(Fcur, w, w, Ic, d, scr);
            for (int e = tid; e < w * w; e += MS_T) {       // D_N -= F_i F_i^T
                const int c = e / w, r = e % w;
                if (r < c) continue;
                double acc = 0.0;
                for (int k = 0; k < d; k++) acc += Fcur[r + k * w] * Fcur[c + k * w];
                DN[e] -= acc;
            }
        }
        for (int e = tid; e < d * d; e += MS_T) { fac[s.offD[i] + e] = Dc[e]; Linv[s.offI[i] + e] = Ic[e]; }
        for (int e = tid; e < o * d; e += MS_T) fac[s.offB[i] + e] = Ccur[e];
        for (int e = tid; e < w * d; e += MS_T) fac[s.offE[i] + e] = Fcur[e];
        double* t = Cprev; Cprev = Ccur; Ccur = t;
        t = Fprev; Fprev = Fcur; Fcur = t;
        __syncthreads();
    }
    if (w > 0) {
        ms_potrf_inv(DN, Dc, rinv, w);     // w <= dmax is not guaranteed: Dc/Ic are sized max(dm, w)^2 by the host
        for (int e = tid; e < w * w; e += MS_T) { fac[s.offD[N - 1] + e] = DN[e]; Linv[s.offI[N - 1] + e] = Dc[e]; }
    }
}

// solve_llt_in_place (:1709-1816) with the inverted pivot blocks: every stage is two small mat-vecs
__global__ void __launch_bounds__(MS_T) ms_solve_kernel(MsDev s, const double* fac_all, const double* inv_all, double* X, const int* active) {
    extern __shared__ __align__(16) double xs[];   // n + dmax scratch
    const int b = blockIdx.x;
    if (active && !active[b]) return;
    const double* fac = fac_all + (size_t)b * s.total;
    const double* Linv = inv_all + (size_t)b * s.total_inv;
    double* x = X + (size_t)b * s.n;
    const int tid = threadIdx.x, N = s.N, w = s.w, n = s.n;
    double* tmp = xs + n;
    for (int i = tid; i < n; i += MS_T) xs[i] = x[i];
    __syncthreads();
    // forward
    for (int i = 0; i + 1 < N; i++) {
        const int d = s.diag[i], st = s.start[i];
        const int po = i > 0 ? s.off[i - 1] : 0, pd = i > 0 ? s.diag[i - 1] : 0, pst = i > 0 ? s.start[i - 1] : 0;
        if (tid < d) {
            double v = xs[st + tid];
            if (tid < po) { const double* C = fac + s.offB[i - 1]; double acc = 0.0; for (int k = 0; k < pd; k++) acc += C[tid + k * po] * xs[pst + k]; v -= acc; }
            tmp[tid] = v;
        }
        __syncthreads();
        if (tid < d) { const double* I = Linv + s.offI[i]; double acc = 0.0; for (int k = 0; k <= tid; k++) acc += I[tid + k * d] * tmp[k]; xs[st + tid] = acc; }
        __syncthreads();
    }
    if (w > 0) {
        if (tid < w) {
            double v = xs[n - w + tid];
            for (int i = 0; i + 1 < N; i++) { const double* F = fac + s.offE[i]; const int d = s.diag[i], st = s.start[i]; double acc = 0.0; for (int k = 0; k < d; k++) acc += F[tid + k * w] * xs[st + k]; v -= acc; }
            tmp[tid] = v;
        }
        __syncthreads();
        const double* I = Linv + s.offI[N - 1];
        double y = 0.0;
        if (tid < w) for (int k = 0; k <= tid; k++) y += I[tid + k * w] * tmp[k];     // L_N^{-1}
        __syncthreads();
        if (tid < w) tmp[tid] = y;
        __syncthreads();
        if (tid < w) { double acc = 0.0; for (int k = tid; k < w; k++) acc += I[k + tid * w] * tmp[k]; xs[n - w + tid] = acc; }   // L_N^{-T}
        __syncthreads();
    }
    // backward
    for (int i = N - 2; i >= 0; i--) {
        const int d = s.diag[i], st = s.start[i], o = (i + 2 < N) ? s.off[i] : 0, nst = (i + 2 < N) ? s.start[i + 1] : 0;
        if (tid < d) {
            double v = xs[st + tid];
            if (o > 0) { const double* C = fac + s.offB[i]; double acc = 0.0; for (int r = 0; r < o; r++) acc += C[r + tid * o] * xs[nst + r]; v -= acc; }
            if (w > 0) { const double* F = fac + s.offE[i]; double acc = 0.0; for (int r = 0; r < w; r++) acc += F[r + tid * w] * xs[n - w + r]; v -= acc; }
            tmp[tid] = v;
        }
        __syncthreads();
        if (tid < d) { const double* I = Linv + s.offI[i]; double acc = 0.0; for (int k = tid; k < d; k++) acc += I[k + tid * d] * tmp[k]; xs[st + tid] = acc; }
        __syncthreads();
    }
    for (int i = tid; i < n; i += MS_T) x[i] = xs[i];
}

__global__ void ms_inv_copy_kernel(const double* z_reg, double* zinv, size_t nz, const double* delta_in, double* delta_out, int batch) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nz) zinv[i] = 1.0 / z_reg[i];
    if (i < (size_t)batch) delta_out[i] = delta_in[i];
}
__global__ void ms_set_ok_kernel(const int* active, int* ok, int batch) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < batch && (!active || active[b])) ok[b] = 1;     // the reference never reports failure (:218)
}
__global__ void ms_copy_masked_kernel(const double* src, double* dst, int len, const int* active) {
    const int b = blockIdx.y;
    if (active && !active[b]) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < len) dst[(size_t)b * len + i] = src[(size_t)b * len + i];
}

// =====================================================================================================
// MultistageBatchedKKT
// =====================================================================================================
static void upload(DevBuf<int>& d, const std::vector<int>& h) {
    d.alloc(std::max<size_t>(h.size(), 1));
    if (!h.empty()) B200_CUDA(cudaMemcpy(d.get(), h.data(), h.size() * sizeof(int), cudaMemcpyHostToDevice));
}

MultistageBatchedKKT::MultistageBatchedKKT(SparseData* data, cudaStream_t st) : D(data) {
    batch = D->batch; n = D->n; p = D->p; m = D->m; stream = st;
    if (!S.detect(D->P, D->AT, D->GT)) throw std::runtime_error(S.error);
    if (S.dmax > MS_T || S.w > MS_T) throw std::runtime_error("multistage: block sizes above 128 are not supported by this build");
    std::vector<int> meta;
    for (auto f : {0, 1, 2}) for (int i = 0; i < S.N; i++) meta.push_back(f == 0 ? S.bi[i].start : f == 1 ? S.bi[i].diag : S.bi[i].off);
    meta.insert(meta.end(), S.offD.begin(), S.offD.end()); meta.insert(meta.end(), S.offB.begin(), S.offB.end());
    meta.insert(meta.end(), S.offE.begin(), S.offE.end()); meta.insert(meta.end(), S.offI.begin(), S.offI.end());
    // warp-chain fast path (multistage_chain.cuh): every front d + o + w fits one warp; solve packets per stage and direction
    warp_chain = S.w <= 32;
    int max_rows = S.w;
    for (int i = 0; i + 1 < S.N; i++) max_rows = std::max(max_rows, S.bi[i].diag + ((i + 2 < S.N) ? S.bi[i].off : 0) + S.w);
    if (max_rows > 32) warp_chain = false;
    if (const char* e = getenv("B200_MS_GENERIC")) if (e[0] == '1') warp_chain = false;
    std::vector<int> cls(S.N, 0), pkF(S.N, 0), szF(S.N, 0), pkB(S.N, 0), szB(S.N, 0);
    size_t slot = 0;
    pk_stride = 0;
    if (warp_chain) {
        auto r2 = [](int v) { return (v + 1) & ~1; };
        for (int i = 0; i < S.N; i++) { const int d = S.bi[i].diag; cls[i] = d <= 8 ? 8 : (d <= 16 ? 16 : 32); }
        plan_partition(cls);
        for (int i = 0; i + 1 < S.N; i++) {
            const int D = cls[i], PD = i > 0 ? cls[i - 1] : 0, ND = (i + 2 < S.N) ? cls[i + 1] : 0;
            szF[i] = D * D + D * PD + r2(S.w * D);
            szB[i] = D * D + D * ND + D * r2(S.w);
            if (part_K > 1 && part_dsep[i] > 0) { szF[i] += part_dsep[i] * D; szB[i] += D * part_dsep[i]; }      // spike blocks Y^T / Y (multistage_partition.cuh)
            pkF[i] = (int)pk_stride; pk_stride += szF[i];
            pkB[i] = (int)pk_stride; pk_stride += szB[i];
            slot = std::max(slot, (size_t)std::max(szF[i], szB[i]));
        }
        chain_rp = (max_rows + 1) / 2;
        chain_slot = (int)slot;
        chain_solve_smem = sizeof(double) * ((size_t)((n + 1) & ~1) + 96 + (size_t)(((MS_META * S.N + 1) / 2 + 1) / 2 * 2) + (size_t)MSW_R * chain_slot);
        if (chain_solve_smem > 227 * 1024 || (size_t)MS_META * S.N * sizeof(int) > 100 * 1024) { warp_chain = false; part_K = 1; }
    }
    for (auto* v : {&cls, &pkF, &szF, &pkB, &szB}) meta.insert(meta.end(), v->begin(), v->end());
    upload(d_meta, meta);
    if (warp_chain && part_K > 1) build_partition(cls, st);
    if (warp_chain) {
        packets.alloc((size_t)batch * pk_stride);
        packets.zero(st);
        allow_dynamic_smem(msw_solve_kernel, (size_t)((int)std::max<size_t>(chain_solve_smem, 48 * 1024)));
        const int csm = (int)(sizeof(MswChainSmem) + sizeof(int) * MS_META * std::max(S.N, MSP_META_MAX));      // segment mode stages MSP_META_MAX entries per array
        allow_dynamic_smem(msw_factor_chain_kernel<4>, (size_t)(csm));
        allow_dynamic_smem(msw_factor_chain_kernel<8>, (size_t)(csm));
        allow_dynamic_smem(msw_factor_chain_kernel<12>, (size_t)(csm));
        allow_dynamic_smem(msw_factor_chain_kernel<14>, (size_t)(csm));
        allow_dynamic_smem(msw_factor_chain_kernel<16>, (size_t)(csm));
    }
    upload(d_P_slot, S.P_slot);
    std::vector<int> diag_var(S.total, -1);
    for (int i = 0; i < n; i++) diag_var[S.diag_slot[i]] = i;
    upload(d_diag_var, diag_var);
    upload(d_a_ptr, S.a_ptr); upload(d_a_qa, S.a_qa); upload(d_a_qb, S.a_qb);
    upload(d_g_ptr, S.g_ptr); upload(d_g_qa, S.g_qa); upload(d_g_qb, S.g_qb); upload(d_g_row, S.g_row);
    const size_t T = (size_t)batch * S.total;
    Pblk.alloc(T); AtAblk.alloc(T); fac.alloc(T); Linv.alloc((size_t)batch * S.total_inv);
    Pblk.zero(st); AtAblk.zero(st); fac.zero(st); Linv.zero(st);
    zinv.alloc(std::max<size_t>((size_t)batch * m, 1)); delta.alloc(batch); work_z.alloc(std::max<size_t>((size_t)batch * m, 1));
    const int dm = std::max(S.dmax, S.w), om = S.omax, w = S.w;
    // the factor kernel's smem carve-up uses dmax := max(dmax, w) so that the arrow corner fits in Dc / Ic
    factor_smem = sizeof(double) * ((size_t)2 * dm * dm + 2 * (size_t)om * dm + 2 * (size_t)w * dm + (size_t)w * w + (size_t)std::max(om, w) * dm + std::max(dm, w) + 8);
    solve_smem = sizeof(double) * ((size_t)n + std::max(dm, w) + 8);
    if (factor_smem > 227 * 1024 || solve_smem > 227 * 1024) throw std::runtime_error("multistage: blocks too large for shared memory");
    allow_dynamic_smem(ms_factor_kernel, (size_t)((int)std::max<size_t>(factor_smem, 48 * 1024)));
    allow_dynamic_smem(ms_solve_kernel, (size_t)((int)std::max<size_t>(solve_smem, 48 * 1024)));
    load_P();
    compute_AtA();
}

static MsDev make_dev(const MsStructure& S, const int* meta) {
    MsDev d;
    const int N = S.N;
    d.start = meta; d.diag = meta + N; d.off = meta + 2 * N; d.offD = meta + 3 * N; d.offB = meta + 4 * N; d.offE = meta + 5 * N; d.offI = meta + 6 * N;
    d.cls = meta + 7 * N; d.pkF = meta + 8 * N; d.szF = meta + 9 * N; d.pkB = meta + 10 * N; d.szB = meta + 11 * N;
    d.N = N; d.w = S.w; d.n = S.n; d.total = S.total; d.total_inv = S.total_inv; d.dmax = std::max(S.dmax, S.w); d.omax = S.omax;
    return d;
}

void MultistageBatchedKKT::load_P() {
    Pblk.zero(stream);
    if (D->P.nnz) { dim3 g(ceil_div(D->P.nnz, 256), batch);
        B200_LAUNCH(ms_scatter_P_kernel, g, 256, 0, stream, d_P_slot.get(), D->P.nnz, S.total, D->Px.get(), Pblk.get()); }
}
void MultistageBatchedKKT::compute_AtA() {
    dim3 g(ceil_div(S.total, 128), batch);
    B200_LAUNCH(ms_accumulate_kernel, g, 128, 0, stream, d_a_ptr.get(), d_a_qa.get(), d_a_qb.get(), (const int*)nullptr, S.total, D->AT.nnz, p, D->ATx.get(),
                (const double*)nullptr, AtAblk.get(), (const int*)nullptr);
}
void MultistageBatchedKKT::update_data(int options) {   // :140-178
    if (options & 1) load_P();
    if (options & 2) compute_AtA();
}
void MultistageBatchedKKT::copy_from(const MultistageBatchedKKT& o) {
    auto cp = [&](DevBuf<double>& d, const DevBuf<double>& s) { if (s.n) B200_CUDA(cudaMemcpyAsync(d.get(), s.get(), s.n * sizeof(double), cudaMemcpyDeviceToDevice, stream)); };
    cp(Pblk, o.Pblk); cp(AtAblk, o.AtAblk); cp(fac, o.fac); cp(Linv, o.Linv); cp(zinv, o.zinv); cp(delta, o.delta); cp(packets, o.packets);
    if (part_K > 1 && o.part_K == part_K) { cp(rfac, o.rfac); cp(rpackets, o.rpackets); }
}

void MultistageBatchedKKT::factor(const double* delta_in, const double* x_reg, const double* z_reg, const int* active, int* ok) {   // :180-219
    B200_ZONE("piqp::MultistageKKT::update_scalings_and_factor");
    const size_t nz = (size_t)batch * m, tot = std::max(nz, (size_t)batch);
    B200_LAUNCH(ms_inv_copy_kernel, (unsigned)((tot + 255) / 256), 256, 0, stream, z_reg, zinv.get(), nz, delta_in, delta.get(), batch);
    tic(T_ASSEMBLE);
    dim3 g(ceil_div(S.total, 128), batch);
    B200_LAUNCH(ms_assemble_kernel, g, 128, 0, stream, d_g_ptr.get(), d_g_qa.get(), d_g_qb.get(), d_g_row.get(), d_diag_var.get(), S.total, D->GT.nnz, m, n,
                Pblk.get(), AtAblk.get(), D->GTx.get(), zinv.get(), delta.get(), x_reg, fac.get(), active);
    toc(T_ASSEMBLE);
    tic(T_FACTOR);
    if (warp_chain) {
        const MsDev dv = make_dev(S, d_meta.get());
        const size_t msm = sizeof(int) * MS_META * S.N;
#define MSW_FACTOR(RP) B200_LAUNCH(msw_factor_kernel<RP>, batch, 64, msm, stream, dv, fac.get(), Linv.get(), packets.get(), pk_stride, active)
#define MSW_CHAIN(RP) B200_LAUNCH(msw_factor_chain_kernel<RP>, batch, 64, sizeof(MswChainSmem) + msm, stream, dv, fac.get(), packets.get(), pk_stride, active, (const int*)nullptr, (double*)nullptr)
        if (part_K > 1) factor_partitioned(dv, active);
        else if (S.w == 0) { if (chain_rp <= 4) MSW_CHAIN(4); else if (chain_rp <= 8) MSW_CHAIN(8); else if (chain_rp <= 12) MSW_CHAIN(12);
                        else if (chain_rp <= 14) MSW_CHAIN(14); else MSW_CHAIN(16); }
        else if (chain_rp <= 4) MSW_FACTOR(4); else if (chain_rp <= 8) MSW_FACTOR(8); else if (chain_rp <= 12) MSW_FACTOR(12);
        else if (chain_rp <= 14) MSW_FACTOR(14); else MSW_FACTOR(16);
#undef MSW_FACTOR
#undef MSW_CHAIN
    } else B200_LAUNCH(ms_factor_kernel, batch, MS_T, factor_smem, stream, make_dev(S, d_meta.get()), fac.get(), Linv.get(), active);
    toc(T_FACTOR);
    B200_LAUNCH(ms_set_ok_kernel, ceil_div(batch, 256), 256, 0, stream, active, ok, batch);
}

void MultistageBatchedKKT::solve(const double* rx, const double* ry, const double* rz, double* lx, double* ly, double* lz, const int* active) {   // :221-288
    B200_ZONE("piqp::MultistageKKT::solve");
    if (n == 0) return;
    tic(T_SOLVE);
    dim3 gn(ceil_div(n, 256), batch);
    B200_LAUNCH(ms_copy_masked_kernel, gn, 256, 0, stream, rx, lx, n, active);
    if (m > 0) spmv_rows(D->GT, D->GTx.get(), 1.0, rz, m, lx, 1, zinv.get(), nullptr, 0, batch, active, stream);          // lx += GT (zinv .* rz)
    if (p > 0) spmv_rows(D->AT, D->ATx.get(), 1.0, ry, p, lx, 1, nullptr, delta.get(), 1, batch, active, stream);         // lx += AT ry / delta
    if (warp_chain && part_K > 1) solve_partitioned(lx, active);
    else if (warp_chain) B200_LAUNCH(msw_solve_kernel, batch, 32, chain_solve_smem, stream, make_dev(S, d_meta.get()), chain_slot, Linv.get(), packets.get(), pk_stride, lx, active);
    else B200_LAUNCH(ms_solve_kernel, batch, MS_T, solve_smem, stream, make_dev(S, d_meta.get()), fac.get(), Linv.get(), lx, active);
    if (p > 0) spmv_cols(D->AT, D->ATx.get(), 1.0, lx, n, ly, ry, 1.0, nullptr, delta.get(), 1, batch, active, stream);    // ly = (A lx - ry) / delta
    if (m > 0) spmv_cols(D->GT, D->GTx.get(), 1.0, lx, n, lz, rz, 1.0, zinv.get(), nullptr, 0, batch, active, stream);     // lz = zinv .* (G lx - rz)
    toc(T_SOLVE);
}

// =====================================================================================================
// parallel-in-horizon partition (multistage_partition.cuh)
// =====================================================================================================
// Chooses K-1 separator stages.  Cost model (measured stage latencies: factor chain 3.1 us, spike / solve stage ~0.6 us):
// factor ~ (N/K) (3.1 + 0.8) + 3.1 K, solve ~ 2 (N/K) 0.65 + 2 K 0.5  =>  K ~ sqrt(1.25 N).
void MultistageBatchedKKT::plan_partition(const std::vector<int>& cls) {
    part_K = 1; part_dsep.assign(S.N, 0); part_bounds.clear(); part_sep.clear();
    const int nreal = S.N - 1;
    if (!warp_chain || S.w != 0 || nreal < 16) return;
    if (const char* e = getenv("B200_MS_NO_PARTITION")) if (e[0] == '1') return;
    int K = (int)std::lround(std::sqrt(1.25 * nreal));
    // The runs of all QPs must be RESIDENT together (a second wave of CTAs would double the chain again): one run's CTA of the
    // factor kernel holds sizeof(MswChainSmem) = 42 KB of shared memory => 5 per SM.  BASELINE config 4 (batch 128 per GPU): K = 5.
    {
        int dev = 0, sms = 148;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int per_sm = std::max(1, (int)((227 * 1024) / (sizeof(MswChainSmem) + sizeof(int) * MS_META * MSP_META_MAX + 1024)));
        K = std::min(K, (sms * per_sm) / std::max(batch, 1));
    }
    if (const char* e = getenv("B200_MS_SEGMENTS")) K = atoi(e);
    K = std::max(1, std::min(K, std::min(32, nreal / 3)));
    if (K < 2) return;
    std::vector<int> sep;
    for (int k = 1; k < K; k++) sep.push_back((int)((long long)k * nreal / K));          // stages 1 .. nreal-2, at least one interior stage in between
    for (size_t k = 0; k < sep.size(); k++) {
        const int lo = k == 0 ? 0 : sep[k - 1] + 2;
        if (sep[k] < std::max(lo, 1) || sep[k] > nreal - 2) return;
    }
    // every reduced front (separator k + coupling rows of separator k+1) must fit one warp like the original fronts
    for (size_t k = 0; k < sep.size(); k++) {
        const int d = S.bi[sep[k]].diag, o = k + 1 < sep.size() ? S.bi[sep[k + 1] - 1].off : 0;
        if (d + o > 32 || S.bi[sep[k] - 1].off > d) return;
    }
    part_K = K; part_sep = sep;
    for (int r = 0; r < K; r++) {
        const int i0 = r == 0 ? 0 : sep[r - 1] + 1, i1 = r + 1 < K ? sep[r] : nreal;
        part_bounds.push_back(i0); part_bounds.push_back(i1);
        if (r > 0) for (int i = i0; i < i1; i++) part_dsep[i] = cls[sep[r - 1]];
    }
}

void MultistageBatchedKKT::build_partition(const std::vector<int>& cls, cudaStream_t st) {
    const int K = part_K, NR = K;                        // reduced chain: K-1 stages + the (empty) arrow entry, like S.bi
    auto r2 = [](int v) { return (v + 1) & ~1; };
    std::vector<int> rs(NR, 0), rd(NR, 0), ro(NR, 0), rD(NR, 0), rB(NR, 0), rE(NR, 0), rI(NR, 0), rc(NR, 8), rpF(NR, 0), rsF(NR, 0), rpB(NR, 0), rsB(NR, 0);
    int pos = 0, tot = 0, toti = 0, rmax = 0;
    for (int k = 0; k < K - 1; k++) {
        const int d = S.bi[part_sep[k]].diag, o = k + 1 < K - 1 ? S.bi[part_sep[k + 1] - 1].off : 0;
        rs[k] = pos; rd[k] = d; ro[k] = o; pos += d;
        rD[k] = tot; tot += d * d; rB[k] = tot; tot += o * d; rE[k] = tot; rI[k] = toti; toti += d * d;
        rc[k] = cls[part_sep[k]];
        rmax = std::max(rmax, d + o);
    }
    rs[NR - 1] = pos; part_rn = pos; part_rtotal = std::max(tot, 1);
    size_t rstride = 0, rslot = 0;
    for (int k = 0; k < K - 1; k++) {
        const int D = rc[k], PD = k > 0 ? rc[k - 1] : 0, ND = (k + 2 < NR) ? rc[k + 1] : 0;
        rsF[k] = D * D + D * PD; rsB[k] = D * D + D * ND;
        rpF[k] = (int)rstride; rstride += rsF[k]; rpB[k] = (int)rstride; rstride += rsB[k];
        rslot = std::max(rslot, (size_t)std::max(rsF[k], rsB[k]));
    }
    part_rpk_stride = std::max<size_t>(rstride, 2); part_rslot = (int)rslot; part_rrp = (rmax + 1) / 2;
    std::vector<int> meta;
    for (auto* v : {&rs, &rd, &ro, &rD, &rB, &rE, &rI, &rc, &rpF, &rsF, &rpB, &rsB}) meta.insert(meta.end(), v->begin(), v->end());
    upload(d_rmeta, meta);
    std::vector<int> aux(part_bounds);                    // seg_bounds[2K] | sep[K-1] | rstart[K] | roffD[K-1] | roffB[K-1]
    aux.insert(aux.end(), part_sep.begin(), part_sep.end());
    aux.insert(aux.end(), rs.begin(), rs.end());
    aux.insert(aux.end(), rD.begin(), rD.begin() + (K - 1));
    aux.insert(aux.end(), rB.begin(), rB.begin() + (K - 1));
    upload(d_part, aux);
    part_seg_len = 0; part_dsep_max = 8;
    for (int r = 0; r < K; r++) {
        const int i0 = part_bounds[2 * r], i1 = part_bounds[2 * r + 1];
        part_seg_len = std::max(part_seg_len, S.bi[i1 - 1].start + S.bi[i1 - 1].diag - S.bi[i0].start);
        if (r > 0) part_dsep_max = std::max(part_dsep_max, cls[part_sep[r - 1]]);
    }
    part_seg_len = r2(part_seg_len);
    rfac.alloc((size_t)batch * part_rtotal); rfac.zero(st);
    rpackets.alloc((size_t)batch * part_rpk_stride); rpackets.zero(st);
    carry.alloc((size_t)batch * K * 1024); carry.zero(st);
    zbuf.alloc((size_t)batch * K * 32); zbuf.zero(st);
    xred.alloc((size_t)batch * std::max(part_rn, 1)); xred.zero(st);
    const size_t meta_d = (size_t)(MS_META * MSP_META_MAX + 1) / 2;
    // ring slots are sized for the MOST COMMON packet size; the few larger packets (stages of a bigger class) are read from global memory
    auto mode_of = [](std::vector<int> v) { std::sort(v.begin(), v.end()); int best = v.empty() ? 2 : v[0], cnt = 0, bc = 0, prev = -1;
        for (int x : v) { cnt = (x == prev) ? cnt + 1 : 1; prev = x; if (cnt >= bc) { bc = cnt; best = x; } } return best; };
    std::vector<int> sz_solve, sz_spike;
    {
        std::vector<int> szF_(S.N, 0), szB_(S.N, 0);
        for (int i = 0; i + 1 < S.N; i++) {
            const int D = cls[i], PD = i > 0 ? cls[i - 1] : 0, ND = (i + 2 < S.N) ? cls[i + 1] : 0;
            const int y = part_dsep[i] * D;
            sz_solve.push_back(std::max(D * D + D * PD + y, D * D + D * ND + y));
            if (i > 0) sz_spike.push_back(D * D + D * PD);
        }
    }
    part_solve_slot = mode_of(sz_solve); part_spike_slot = mode_of(sz_spike);
    part_seg_smem = sizeof(double) * ((size_t)part_seg_len + 96 + 64 + meta_d + (size_t)MSP_PF * part_solve_slot);
    part_spike_smem = sizeof(double) * ((size_t)3 * 32 * MSP_LDY + meta_d + (size_t)MSP_PF * part_spike_slot);
    for (int r = 0; r < K; r++) if (part_bounds[2 * r + 1] - part_bounds[2 * r] + 2 > MSP_META_MAX) { part_K = 1; return; }      // run-local meta copies
    part_rsolve_smem = sizeof(double) * ((size_t)((part_rn + 1) & ~1) + 96 + (size_t)(((MS_META * NR + 1) / 2 + 1) / 2 * 2) + (size_t)MSW_R * part_rslot);
    if (part_seg_smem > 227 * 1024 || part_spike_smem > 227 * 1024 || part_rsolve_smem > 227 * 1024) { part_K = 1; return; }
    allow_dynamic_smem(msp_fwd_kernel, (size_t)((int)std::max<size_t>(part_seg_smem, 48 * 1024)));
    allow_dynamic_smem(msp_bwd_kernel, (size_t)((int)std::max<size_t>(part_seg_smem, 48 * 1024)));
    allow_dynamic_smem(msp_spike_kernel, (size_t)((int)std::max<size_t>(part_spike_smem, 48 * 1024)));
    // fused solve (one CTA per QP, 2K warps): worth it while the QPs of the batch are resident together
    part_fused_smem = 0;
    {
        int dev = 0, sms = 148;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        part_fused_slot = part_spike_slot;              // chain part of the packets: inv(L) | (B L^-T), the same mode as the spike kernel's
        size_t run_d = (size_t)part_seg_len + 96 + 96 + (size_t)MSF_PF * part_fused_slot + (MS_META * MSP_META_MAX + 2 + 1) / 2;
        run_d = (run_d + 1) & ~(size_t)1;
        const size_t total = sizeof(double) * (K * run_d) + part_rsolve_smem;
        const char* e = getenv("B200_MS_FUSED_SOLVE");
        const bool want = e ? e[0] == '1' : batch <= 2 * sms;
        if (want && 2 * K * 32 <= 512 && total <= 200 * 1024) {
            part_fused_smem = total; part_fused_run = (int)run_d;
            allow_dynamic_smem(msp_solve_fused_kernel, total);
        }
    }
}

static MsDev make_rdev(const int* meta, int NR, int rn, int rtotal) {
    MsDev d;
    d.start = meta; d.diag = meta + NR; d.off = meta + 2 * NR; d.offD = meta + 3 * NR; d.offB = meta + 4 * NR; d.offE = meta + 5 * NR; d.offI = meta + 6 * NR;
    d.cls = meta + 7 * NR; d.pkF = meta + 8 * NR; d.szF = meta + 9 * NR; d.pkB = meta + 10 * NR; d.szB = meta + 11 * NR;
    d.N = NR; d.w = 0; d.n = rn; d.total = rtotal; d.total_inv = 0; d.dmax = 32; d.omax = 32;
    return d;
}
MsPart MultistageBatchedKKT::make_part() const {
    MsPart P;
    const int K = part_K;
    const int* a = d_part.get();
    P.seg_bounds = a; P.sep = a + 2 * K; P.rstart = a + 2 * K + (K - 1); P.roffD = a + 2 * K + (K - 1) + K; P.roffB = P.roffD + (K - 1);
    P.K = K; P.rn = part_rn; P.rtotal = part_rtotal;
    return P;
}

// B200_MS_TIMING=1: CUDA events between the launches of the partitioned factor / solve (in-situ device times per kernel, warm caches;
// adds one stream synchronisation per call, so only for diagnostics)
struct MsPartTimer {
    static constexpr int MAXE = 8;
    cudaEvent_t ev[MAXE]; double acc[2][MAXE] = {}; long calls[2] = {0, 0}; bool on = false;
    MsPartTimer() { on = getenv("B200_MS_TIMING") != nullptr; if (on) for (auto& e : ev) cudaEventCreate(&e); }
    ~MsPartTimer() {
        if (!on) return;
        const char* nf[] = {"chain(runs)", "spike", "reduce_assemble", "chain(reduced)"};
        const char* ns[] = {"fwd(runs)", "gather", "solve(reduced)", "bwd(runs)"};
        for (int w = 0; w < 2; w++) { if (!calls[w]) continue; fprintf(stderr, "[B200_MS_TIMING] %s, %ld calls, us per call:", w ? "solve" : "factor", calls[w]);
            for (int k = 0; k < 4; k++) fprintf(stderr, "  %s %.1f", w ? ns[k] : nf[k], 1e3 * acc[w][k] / calls[w]); fprintf(stderr, "\n"); }
    }
    void mark(int k, cudaStream_t st) { if (on) cudaEventRecord(ev[k], st); }
    void done(int which, int n, cudaStream_t st) {
        if (!on) return;
        cudaStreamSynchronize(st);
        for (int k = 0; k < n; k++) { float ms = 0; cudaEventElapsedTime(&ms, ev[k], ev[k + 1]); acc[which][k] += ms; }
        calls[which]++;
    }
};
static MsPartTimer g_ms_timer;
bool MultistageBatchedKKT::graph_capturable() const { return !g_ms_timer.on; }

void MultistageBatchedKKT::factor_partitioned(const MsDev& dv, const int* active) {
    B200_ZONE("piqp::MultistageKKT::factor_kkt");
    const int K = part_K;
    const MsPart P = make_part();
    dim3 gseg(batch, K);
    g_ms_timer.mark(0, stream);
#define MSP_CHAIN(RP) B200_LAUNCH(msw_factor_chain_kernel<RP>, gseg, 64, sizeof(MswChainSmem) + sizeof(int) * MS_META * MSP_META_MAX, stream, dv, fac.get(), packets.get(), pk_stride, active, P.seg_bounds, carry.get())
    if (chain_rp <= 4) MSP_CHAIN(4); else if (chain_rp <= 8) MSP_CHAIN(8); else if (chain_rp <= 12) MSP_CHAIN(12); else if (chain_rp <= 14) MSP_CHAIN(14); else MSP_CHAIN(16);
#undef MSP_CHAIN
    g_ms_timer.mark(1, stream);
    B200_LAUNCH(msp_spike_kernel, dim3(batch, K - 1), 32, part_spike_smem, stream, dv, P, part_spike_slot, fac.get(), packets.get(), pk_stride, carry.get(), rfac.get(), active);
    g_ms_timer.mark(2, stream);
    g_ms_timer.mark(3, stream);
    const MsDev rd = make_rdev(d_rmeta.get(), K, part_rn, part_rtotal);
    const size_t rmsm = sizeof(int) * MS_META * K;
#define MSP_RCHAIN(RP) B200_LAUNCH(msw_factor_chain_kernel<RP>, batch, 64, sizeof(MswChainSmem) + rmsm, stream, rd, rfac.get(), rpackets.get(), part_rpk_stride, active, (const int*)nullptr, (double*)nullptr)
    if (part_rrp <= 4) MSP_RCHAIN(4); else if (part_rrp <= 8) MSP_RCHAIN(8); else if (part_rrp <= 12) MSP_RCHAIN(12); else if (part_rrp <= 14) MSP_RCHAIN(14); else MSP_RCHAIN(16);
#undef MSP_RCHAIN
    g_ms_timer.mark(4, stream);
    g_ms_timer.done(0, 4, stream);
}

void MultistageBatchedKKT::solve_partitioned(double* lx, const int* active) {
    B200_ZONE("piqp::MultistageKKT::solve_llt_in_place");
    const int K = part_K;
    const MsPart P = make_part();
    const MsDev dv = make_dev(S, d_meta.get());
    dim3 gseg(batch, K);
    if (part_fused_smem > 0 && !g_ms_timer.on) {
        const MsDev rd = make_rdev(d_rmeta.get(), K, part_rn, part_rtotal);
        B200_LAUNCH(msp_solve_fused_kernel, batch, 64 * K, part_fused_smem, stream, dv, P, rd, part_fused_slot, part_seg_len, part_fused_run, part_rslot, packets.get(), pk_stride,
                    rpackets.get(), part_rpk_stride, lx, zbuf.get(), xred.get(), active);
        return;
    }
    g_ms_timer.mark(0, stream);
    B200_LAUNCH(msp_fwd_kernel, gseg, 32, part_seg_smem, stream, dv, P, part_solve_slot, part_seg_len, packets.get(), pk_stride, lx, zbuf.get(), active);
    g_ms_timer.mark(1, stream);
    B200_LAUNCH(msp_gather_kernel, dim3(batch, ceil_div(K - 1, 4)), 128, 0, stream, dv, P, packets.get(), pk_stride, lx, zbuf.get(), xred.get(), active);
    const MsDev rd = make_rdev(d_rmeta.get(), K, part_rn, part_rtotal);
    g_ms_timer.mark(2, stream);
    B200_LAUNCH(msw_solve_kernel, batch, 32, std::max<size_t>(part_rsolve_smem, 1), stream, rd, part_rslot, (const double*)nullptr, rpackets.get(), part_rpk_stride, xred.get(), active);
    g_ms_timer.mark(3, stream);
    B200_LAUNCH(msp_bwd_kernel, gseg, 32, part_seg_smem, stream, dv, P, part_solve_slot, part_seg_len, packets.get(), pk_stride, lx, xred.get(), active);
    g_ms_timer.mark(4, stream);
    g_ms_timer.done(1, 4, stream);
}

void MultistageBatchedKKT::eval_P_x(double alpha, const double* x, double* z, const int* active) { spmv_sym_upper(D->P, D->Px.get(), alpha, x, z, batch, active, stream); }
void MultistageBatchedKKT::eval_A(double an, double at, const double* xn, const double* xt, double* zn, double* zt, const int* active) {
    if (p > 0) spmv_cols(D->AT, D->ATx.get(), an, xn, n, zn, nullptr, 0.0, nullptr, nullptr, 0, batch, active, stream);
    spmv_rows(D->AT, D->ATx.get(), at, xt, p, zt, 0, nullptr, nullptr, 0, batch, active, stream);
}
void MultistageBatchedKKT::eval_G(double an, double at, const double* xn, const double* xt, double* zn, double* zt, const int* active) {
    if (m > 0) spmv_cols(D->GT, D->GTx.get(), an, xn, n, zn, nullptr, 0.0, nullptr, nullptr, 0, batch, active, stream);
    spmv_rows(D->GT, D->GTx.get(), at, xt, m, zt, 0, nullptr, nullptr, 0, batch, active, stream);
}
void MultistageBatchedKKT::extract_P_diag(double* P_diag) { sparse_extract_diag(*D, P_diag, stream); }
void MultistageBatchedKKT::print_info() const {   // :385-392
    printf("block sizes:");
    for (int i = 0; i + 1 < S.N; i++) printf(" %d,%d", S.bi[i].diag, S.bi[i].off);
    printf("\narrow width: %d\n", S.bi[S.N - 1].diag);
}

}  // namespace b200
