// piqp_b200/csrc/sparse_wide.cuh -- whole-GPU ("wide") schedule of the supernodal multifrontal LDL^T and of the supernodal
// triangular solves, for FEW LARGE QPs (BASELINE config 3: one sparse QP with n_kkt = 20 000 whose root front has ~10 000 rows).
//
// sparse_frontal.cuh gives every QP one CTA, which is the right shape for a batch of small QPs and leaves 147 SMs idle for a
// single large one.  Here the supernodal elimination tree is walked by LEVELS (all supernodes of a level are independent):
//   * fronts that fit shared memory: one CTA per (supernode, QP), mfw_small_kernel (same arithmetic as mf_factor_kernel);
//   * larger fronts live in HBM (column-major, padded ld) and are processed by kernels that span the GPU:
//       mfw_zero / mfw_scatter      original entries
//       mfw_pull_kernel             extend-add of the children, PULL form: one warp per front column walks the children that
//                                   hold this column in a fixed order (deterministic, no atomics)
//       mfw_panel_kernel            per 64-column panel: LDL^T of the 64 x 64 pivot block (every CTA redoes it in shared
//                                   memory: no extra launch, identical bits) + one thread per row of the panel solve
//       gemm_nt_tile_kernel<EPI_SUB,true>  trailing update F22 -= L21 D L21^T on the DMMA tensor pipe (dense backend's kernel)
//       mfw_schur_kernel            Schur complement -> the supernode's own update slot
// Solves (L y = b, D, L^T x = y) use the PULL form as well (row view of L forward, column view backward), so that concurrent
// supernodes never write the same entry: one warp per narrow supernode per level, and for wide supernodes a blocked dense
// solve over the GPU (mfw_fwd_block_kernel / mfw_bwd_block_kernel, one launch per 128-column block).
// Replaces LDLt::factorize_numeric_upper_triangular / solve_inplace (include/piqp/sparse/ldlt.hpp:101-218).
#pragma once
#include "sparse_frontal.cuh"

namespace b200 {

constexpr int MW_NB = 64;      // panel width of the blocked LDL^T of an HBM front
constexpr int MW_T = 256;

__device__ __forceinline__ size_t mfw_colbase(int lp0, int f, int k) {      // L(i, j0 + k) of a supernode sits at Lx[colbase(k) + i], i = front row > k
    return (size_t)((long long)lp0 + (long long)k * (f - 1) - ((long long)k * (k - 1)) / 2 - (k + 1));
}

// ---------------------------------------------------------------------------------------------------------------
// factorisation
// ---------------------------------------------------------------------------------------------------------------
// one CTA per (supernode of this level whose front fits shared memory, QP)
__global__ void __launch_bounds__(MF_T) mfw_small_kernel(MfDev M, const int* __restrict__ list, int fpad, const int* __restrict__ crecw,
                                                         const long long* __restrict__ upd_off_w, const double* __restrict__ PKasm_all,
                                                         double* __restrict__ Lx_all, double* __restrict__ Dv_all, double* __restrict__ Dinv_all,
                                                         double* __restrict__ upd_all, long long upd_stride, int* __restrict__ fail, const int* __restrict__ active) {
    extern __shared__ __align__(16) double mf_sm[];
    const int b = blockIdx.y;
    if (active && !active[b]) return;
    const int s = list[blockIdx.x];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr int NW = MF_T / 32;
    const double* PK = PKasm_all + (size_t)b * M.nnzPK;
    double* Lx = Lx_all + (size_t)b * M.nnzL;
    double* upd = upd_all + (size_t)b * (size_t)upd_stride;
    double* lcol = mf_sm;
    int* relbuf = reinterpret_cast<int*>(mf_sm + fpad);
    double* Fs = mf_sm + ((fpad + fpad / 2 + 3) & ~3);
    const int4 h0 = reinterpret_cast<const int4*>(M.hdr)[2 * s], h1 = reinterpret_cast<const int4*>(M.hdr)[2 * s + 1];
    const int j0 = h0.x, ws = h0.y, us = h0.z, lp0 = h0.w, ab = h1.x, an = h1.y, cb = h1.z, cn = h1.w;
    const int f = ws + us;
    for (int e = tid; e < f * f; e += MF_T) Fs[e] = 0.0;
    __syncthreads();
    for (int t = tid; t < an; t += MF_T) Fs[M.asm_pos[ab + t]] = PK[ab + t];
    for (int c = 0; c < cn; c++) {                       // extend-add, one child at a time (fixed order)
        const int4 cr = reinterpret_cast<const int4*>(crecw)[cb + c];
        const int uc = cr.x;
        const double* U = upd + (((long long)cr.w << 32) | (unsigned)cr.z);
        for (int t = tid; t < uc; t += MF_T) relbuf[t] = M.rel_idx[cr.y + t];
        __syncthreads();
        const int total = uc * uc;
        for (int e = tid; e < total; e += MF_T) {
            const int col = e / uc, a = e - col * uc;
            if (a >= col) Fs[(size_t)relbuf[a] + (size_t)relbuf[col] * f] += U[e];
        }
        __syncthreads();
    }
    if (cn == 0) __syncthreads();
    mf_eliminate_smem(Fs, f, ws, j0, lp0, lcol, Lx, Dv_all + (size_t)b * M.nk, Dinv_all + (size_t)b * M.nk, fail + b);
    if (us > 0) {
        double* U = upd + upd_off_w[s];
        for (int col = wid; col < us; col += NW) {
            const double* Fc = Fs + (size_t)(ws + col) * f + ws;
            for (int a = col + lane; a < us; a += 32) U[a + (size_t)col * us] = Fc[a];
        }
    }
}

__global__ void mfw_zero_kernel(double* __restrict__ F_all, long long stride, long long count, const int* __restrict__ active) {
    const int b = blockIdx.y;
    if (active && !active[b]) return;
    double2* F = reinterpret_cast<double2*>(F_all + (size_t)b * stride);
    const long long n2 = count / 2;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n2; e += (long long)gridDim.x * blockDim.x) F[e] = make_double2(0.0, 0.0);
}
// original entries of the front: asm_pos = row + col * f (front-local), F(row, col) at F[shift + row + col * ld]
__global__ void mfw_scatter_kernel(double* __restrict__ F_all, long long stride, int ld, int shift, int f, const int* __restrict__ asm_pos,
                                   const double* __restrict__ PKasm_all, size_t nnzPK, int ab, int an, const int* __restrict__ active) {
    const int b = blockIdx.y;
    if (active && !active[b]) return;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= an) return;
    const int pos = asm_pos[ab + t], col = pos / f, row = pos - col * f;
    F_all[(size_t)b * stride + shift + row + (size_t)col * ld] = PKasm_all[(size_t)b * nnzPK + ab + t];
}
// extend-add in pull form: warp w owns front column c; pull_ptr[c .. c+1) lists (child record index, column of the child's update
// matrix that maps onto c) in the children's fixed order
__global__ void __launch_bounds__(MW_T) mfw_pull_kernel(double* __restrict__ F_all, long long stride, int ld, int shift, int f, const int* __restrict__ pull_ptr,
                                                        const int* __restrict__ pull_child, const int* __restrict__ pull_cc, const int* __restrict__ crecw,
                                                        const int* __restrict__ rel_idx, const double* __restrict__ upd_all, long long upd_stride,
                                                        const int* __restrict__ active) {
    const int b = blockIdx.y;
    if (active && !active[b]) return;
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x * (MW_T / 32) + (threadIdx.x >> 5);
    if (c >= f) return;
    double* Fc = F_all + (size_t)b * stride + shift + (size_t)c * ld;
    const double* upd = upd_all + (size_t)b * (size_t)upd_stride;
    const int t1 = pull_ptr[c + 1];
    for (int t = pull_ptr[c]; t < t1; t++) {
        const int4 cr = reinterpret_cast<const int4*>(crecw)[pull_child[t]];
        const int cc = pull_cc[t], uc = cr.x;
        const double* U = upd + (((long long)cr.w << 32) | (unsigned)cr.z) + (size_t)cc * uc;
        const int* rel = rel_idx + cr.y;
        for (int a = cc + lane; a < uc; a += 32) Fc[rel[a]] += U[a];
    }
}

// One panel of the blocked LDL^T of an HBM front: columns [k0, k0 + nb) of F (f x f, lower, ld), rows below r0 = k0 + nb.
// Every CTA factors the nb x nb pivot block in shared memory (A11 in F is only read, never overwritten: L11 goes to Lx, D to
// Dv), then each thread solves one row of the panel: w = a L11^-T (= l D), l = w / d; F keeps l (operand of the trailing update).
__global__ void __launch_bounds__(MW_T) mfw_panel_kernel(double* __restrict__ F_all, long long stride, int ld, int shift, int f, int k0, int nb, int j0, int lp0,
                                                         double* __restrict__ Lx_all, size_t nnzL, double* __restrict__ Dv_all, double* __restrict__ Dinv_all, int nk,
                                                         int* __restrict__ fail, const int* __restrict__ active) {
    __shared__ double A[MW_NB][MW_NB + 1];     // lower: A11, overwritten column by column with L11; upper: the unscaled columns w (transposed); diagonal: D
    const int b = blockIdx.y;
    if (active && !active[b]) return;
    const int tid = threadIdx.x;
    double* F = F_all + (size_t)b * stride + shift;
    double* Lx = Lx_all + (size_t)b * nnzL;
    for (int e = tid; e < MW_NB * MW_NB; e += MW_T) {
        const int i = e % MW_NB, c = e / MW_NB;
        if (c <= i) A[i][c] = (i < nb) ? F[(size_t)(k0 + i) + (size_t)(k0 + c) * ld] : 0.0;
    }
    __syncthreads();
    for (int k = 0; k < nb; k++) {
        const double d = A[k][k];
        if (tid > k && tid < nb) { const double w = A[tid][k]; A[k][tid] = w; A[tid][k] = w / d; }
        __syncthreads();
        // A(i, c) -= w_i l_c  for k < c <= i < nb
        const int rem = nb - k - 1;
        for (int e = tid; e < rem * rem; e += MW_T) {
            const int i = k + 1 + e % rem, c = k + 1 + e / rem;
            if (c <= i) A[i][c] -= A[k][i] * A[c][k];
        }
        __syncthreads();
    }
    if (blockIdx.x == 0) {
        for (int k = tid; k < nb; k += MW_T) {
            const double d = A[k][k];
            if (d == 0.0 && fail[b] == 0) fail[b] = j0 + k0 + k + 1;           // ldlt.hpp:161
            Dv_all[(size_t)b * nk + j0 + k0 + k] = d; Dinv_all[(size_t)b * nk + j0 + k0 + k] = 1.0 / d;
        }
        for (int e = tid; e < nb * nb; e += MW_T) {
            const int i = e % nb, c = e / nb;
            if (i > c) Lx[mfw_colbase(lp0, f, k0 + c) + (k0 + i)] = A[i][c];
        }
    }
    const int r0 = k0 + nb;
    const int i = r0 + blockIdx.x * MW_T + tid;
    if (i >= f) return;
    double w[MW_NB];
#pragma unroll
    for (int k = 0; k < MW_NB; k++) w[k] = (k < nb) ? F[(size_t)i + (size_t)(k0 + k) * ld] : 0.0;
#pragma unroll
    for (int k = 0; k < MW_NB; k++) {
        if (k < nb) {
            double acc = w[k];
#pragma unroll
            for (int q = 0; q < k; q++) acc -= w[q] * A[k][q];
            w[k] = acc;
        }
    }
#pragma unroll
    for (int k = 0; k < MW_NB; k++) {
        if (k < nb) {
            const double l = w[k] / A[k][k];
            F[(size_t)i + (size_t)(k0 + k) * ld] = l;
            Lx[mfw_colbase(lp0, f, k0 + k) + i] = l;
        }
    }
}
// Schur complement of an HBM front -> the supernode's update slot (us x us, ld = us, lower part)
__global__ void mfw_schur_kernel(const double* __restrict__ F_all, long long stride, int ld, int shift, int ws, int us, double* __restrict__ upd_all,
                                 long long upd_stride, long long off, const int* __restrict__ active) {
    const int b = blockIdx.z;
    if (active && !active[b]) return;
    const int col = blockIdx.y;
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= us || a < col) return;
    upd_all[(size_t)b * (size_t)upd_stride + off + a + (size_t)col * us] = F_all[(size_t)b * stride + shift + (ws + a) + (size_t)(ws + col) * ld];
}

// ---------------------------------------------------------------------------------------------------------------
// solves (work: permuted vector [batch][nk] in HBM)
// ---------------------------------------------------------------------------------------------------------------
// forward, narrow supernodes of one level: one warp per supernode, rows in order, y_j = b_j - sum_{k<j} L(j,k) y_k through the row view
__global__ void __launch_bounds__(MW_T) mfw_fwd_small_kernel(const int* __restrict__ list, int cnt, const int* __restrict__ hdr, const int* __restrict__ Rp,
                                                             const int* __restrict__ Rcol, const int* __restrict__ Rpos, const double* __restrict__ Lx_all, size_t nnzL,
                                                             double* __restrict__ work_all, int nk, const int* __restrict__ active) {
    const int b = blockIdx.y;
    if (active && !active[b]) return;
    const int lane = threadIdx.x & 31;
    const int t = blockIdx.x * (MW_T / 32) + (threadIdx.x >> 5);
    if (t >= cnt) return;
    const int s = list[t];
    const int j0 = hdr[8 * s], ws = hdr[8 * s + 1];
    const double* Lx = Lx_all + (size_t)b * nnzL;
    double* w = work_all + (size_t)b * nk;
    for (int j = j0; j < j0 + ws; j++) {
        double acc = 0.0;
        for (int q = Rp[j] + lane; q < Rp[j + 1]; q += 32) acc += Lx[Rpos[q]] * w[Rcol[q]];
        acc = warp_sum(acc);
        if (lane == 0) w[j] -= acc;
        __syncwarp();
    }
}
// forward, wide supernode, part 1: contributions of the columns BEFORE the supernode (descendants) to its rows, one warp per row
__global__ void __launch_bounds__(MW_T) mfw_fwd_pull_kernel(int j0, int ws, const int* __restrict__ Rp, const int* __restrict__ Rcol, const int* __restrict__ Rpos,
                                                            const double* __restrict__ Lx_all, size_t nnzL, double* __restrict__ work_all, int nk,
                                                            const int* __restrict__ active) {
    const int b = blockIdx.y;
    if (active && !active[b]) return;
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * (MW_T / 32) + (threadIdx.x >> 5);
    if (r >= ws) return;
    const int j = j0 + r;
    const double* Lx = Lx_all + (size_t)b * nnzL;
    double* w = work_all + (size_t)b * nk;
    double acc = 0.0;
    for (int q = Rp[j] + lane; q < Rp[j + 1]; q += 32) { const int c = Rcol[q]; if (c < j0) acc += Lx[Rpos[q]] * w[c]; }     // row sorted by column: the tail (c >= j0) is the supernode's own triangle
    acc = warp_sum(acc);
    if (lane == 0) w[j] -= acc;
}
// forward, wide supernode, part 2 (one launch per column block): CTA c owns rows [rb + c*sb, +sb) of the supernode's triangle;
// it subtracts L[rows, gc0 .. gc0+gcn) y[gc0 .. gc0+gcn) (final since the previous launch); CTA 0 then solves its own sb x sb
// unit-lower triangle, which makes block rb final for the next launch.  smem: sb*(sb+1) + 3*sb doubles.
__global__ void __launch_bounds__(MW_T) mfw_fwd_block_kernel(int j0, int ws, int f, int lp0, int gc0, int gcn, int sb, const double* __restrict__ Lx_all, size_t nnzL,
                                                             double* __restrict__ work_all, int nk, const int* __restrict__ active) {
    extern __shared__ __align__(16) double sm[];
    const int b = blockIdx.y;
    if (active && !active[b]) return;
    const int tid = threadIdx.x;
    const double* Lx = Lx_all + (size_t)b * nnzL;
    double* w = work_all + (size_t)b * nk + j0;
    double* xs = sm;                 // sb
    double* part = sm + sb;          // 2 * sb  (upper half of the threads)
    double* T = sm + 3 * sb;         // sb x (sb + 1)
    const int rb = gc0 + gcn;
    const int r0 = rb + blockIdx.x * sb, rn = min(sb, ws - r0);
    if (rn <= 0) return;
    const int half = tid / sb, r = tid - half * sb;       // MW_T >= 2 * sb is not required: threads beyond 2*sb idle in the mat-vec
    double acc = 0.0;
    if (half < 2 && r < rn && gcn > 0) {
        const int i = r0 + r;
        const int kh = (gcn + 1) / 2, ka = gc0 + half * kh, kb = min(gc0 + gcn, ka + kh);
        for (int k = ka; k < kb; k++) acc += Lx[mfw_colbase(lp0, f, k) + i] * w[k];
    }
    if (half == 1 && r < rn) part[r] = acc;
    __syncthreads();
    if (half == 0 && r < rn) xs[r] = w[r0 + r] - (acc + part[r]);
    if (blockIdx.x != 0) {
        if (half == 0 && r < rn) w[r0 + r] = xs[r];
        return;
    }
    // CTA 0: unit-lower triangle of block [r0, r0 + rn)
    for (int e = tid; e < rn * rn; e += MW_T) {
        const int i = e % rn, k = e / rn;
        if (i > k) T[i * (sb + 1) + k] = Lx[mfw_colbase(lp0, f, r0 + k) + (r0 + i)];
    }
    __syncthreads();
    for (int k = 0; k + 1 < rn; k++) {
        const double xk = xs[k];
        if (tid > k && tid < rn) xs[tid] -= T[tid * (sb + 1) + k] * xk;
        __syncthreads();
    }
    if (tid < rn) w[r0 + tid] = xs[tid];
}
// backward, wide supernode (one launch per column block [c0, c0 + cn), last block first): CTA k computes
// dot_k = sum_{i >= c0+cn} L(i, c0+k) x_i  (rows of the triangle below the block, then the update rows through their row
// indices); the CTA that finishes last subtracts the dots and solves the cn x cn transposed unit triangle.
// smem: sb*(sb+1) + 2*sb + 32 doubles.
__global__ void __launch_bounds__(MW_T) mfw_bwd_block_kernel(int j0, int ws, int f, int lp0, const int* __restrict__ li_u, int c0, int cn, int sb,
                                                             const double* __restrict__ Lx_all, size_t nnzL, double* __restrict__ work_all, int nk,
                                                             double* __restrict__ tmp_all, unsigned* __restrict__ counter, const int* __restrict__ active) {
    extern __shared__ __align__(16) double sm[];
    __shared__ int is_last;
    const int b = blockIdx.y;
    if (active && !active[b]) return;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const double* Lx = Lx_all + (size_t)b * nnzL;
    double* wg = work_all + (size_t)b * nk;
    double* tmp = tmp_all + (size_t)b * sb;
    double* xs = sm;                 // sb
    double* red = sm + sb;           // 32
    double* T = sm + 2 * sb + 32;    // sb x (sb + 1)
    const int k = blockIdx.x;
    {
        const size_t base = mfw_colbase(lp0, f, c0 + k);
        double acc = 0.0;
        for (int i = c0 + cn + tid; i < f; i += MW_T) {
            const double xi = i < ws ? wg[j0 + i] : wg[li_u[i - ws]];
            acc += Lx[base + i] * xi;
        }
        acc = warp_sum(acc);
        if (lane == 0) red[wid] = acc;
        __syncthreads();
        if (tid == 0) {
            double t2 = 0.0;
            for (int w2 = 0; w2 < MW_T / 32; w2++) t2 += red[w2];
            tmp[k] = t2;
            __threadfence();
            const unsigned prev = atomicAdd(counter + b, 1u);
            is_last = (prev == (unsigned)(cn - 1));
        }
        __syncthreads();
    }
    if (!is_last) return;
    __threadfence();
    for (int e = tid; e < cn * cn; e += MW_T) {
        const int i = e % cn, kk = e / cn;
        if (i > kk) T[i * (sb + 1) + kk] = Lx[mfw_colbase(lp0, f, c0 + kk) + (c0 + i)];
    }
    if (tid < cn) xs[tid] = wg[j0 + c0 + tid] - __ldcg(tmp + tid);
    __syncthreads();
    for (int i = cn - 1; i > 0; i--) {           // x_k -= L(i, k) x_i for k < i, x_i final
        const double xi = xs[i];
        if (tid < i) xs[tid] -= T[i * (sb + 1) + tid] * xi;
        __syncthreads();
    }
    if (tid < cn) wg[j0 + c0 + tid] = xs[tid];
    if (tid == 0) counter[b] = 0;
}
// backward, narrow supernodes of one level: one warp per supernode, columns in reverse, x_j = y_j - sum_{i>j} L(i,j) x_i
__global__ void __launch_bounds__(MW_T) mfw_bwd_small_kernel(const int* __restrict__ list, int cnt, const int* __restrict__ hdr, const int* __restrict__ Lp,
                                                             const int* __restrict__ Li, const double* __restrict__ Lx_all, size_t nnzL, double* __restrict__ work_all,
                                                             int nk, const int* __restrict__ active) {
    const int b = blockIdx.y;
    if (active && !active[b]) return;
    const int lane = threadIdx.x & 31;
    const int t = blockIdx.x * (MW_T / 32) + (threadIdx.x >> 5);
    if (t >= cnt) return;
    const int s = list[t];
    const int j0 = hdr[8 * s], ws = hdr[8 * s + 1];
    const double* Lx = Lx_all + (size_t)b * nnzL;
    double* w = work_all + (size_t)b * nk;
    for (int j = j0 + ws - 1; j >= j0; j--) {
        double acc = 0.0;
        for (int e = Lp[j] + lane; e < Lp[j + 1]; e += 32) acc += Lx[e] * w[Li[e]];
        acc = warp_sum(acc);
        if (lane == 0) w[j] -= acc;
        __syncwarp();
    }
}

}  // namespace b200
