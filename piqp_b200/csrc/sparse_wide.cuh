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
// solve over the GPU (mfw_fwd_block_kernel / mfw_bwd_block_kernel, one launch per 128-column block: a mat-vec with the block's
// columns plus the precomputed inverse of its diagonal block).  The blocked LDL^T is two-level (groups of panels, window and
// far updates) with look-ahead on a second stream: see SparseLdltBatchedKKT::factor_wide.
// Replaces LDLt::factorize_numeric_upper_triangular / solve_inplace (include/piqp/sparse/ldlt.hpp:101-218).
#pragma once
#include "sparse_frontal.cuh"

namespace b200 {

constexpr int MW_NB = 64;      // panel width of the blocked LDL^T of an HBM front
constexpr int MW_T = 256;
constexpr int MW_TS = 1024;    // threads of the blocked-solve kernels: the block steps are latency chains, more threads = fewer dependent loads per thread

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
    int t = pull_ptr[c];
    const int4* rec = reinterpret_cast<const int4*>(crecw);
    int4 cr = make_int4(0, 0, 0, 0); int cc = 0;
    if (t < t1) { cc = pull_cc[t]; cr = rec[pull_child[t]]; }
    for (; t < t1; t++) {
        int4 crn = cr; int ccn = cc;
        if (t + 1 < t1) { ccn = pull_cc[t + 1]; crn = rec[pull_child[t + 1]]; }      // next child's record: off the dependent chain
        const int uc = cr.x;
        const double* U = upd + (((long long)cr.w << 32) | (unsigned)cr.z) + (size_t)cc * uc;
        const int* rel = rel_idx + cr.y;
        for (int a = cc + lane; a < uc; a += 32) Fc[rel[a]] += U[a];
        __syncwarp();                                   // two children may touch the same row from different lanes: keep their order
        cr = crn; cc = ccn;
    }
}

// One panel of the blocked LDL^T of an HBM front: columns [k0, k0 + nb) of F (f x f, lower, ld), rows below r0 = k0 + nb.
// Every CTA factors the nb x nb pivot block itself (A11 in F is only read, never overwritten: L11 goes to Lx, D to Dv): thread
// (i, g) keeps A11(i, c), c = g mod 4, in 16 REGISTERS that are rotated by one slot after every 4 pivots, so that the pivot
// loop stays rolled (the first, fully unrolled version was instruction-fetch bound: 5.8 no-instruction stalls per issue) and
// still indexes registers statically; one barrier per pivot through a double-buffered pivot column.  Then each thread solves
// one row of the panel, l = a L11^-T D^-1, in 16-column chunks: the contributions of earlier chunks come from a private
// shared-memory column (rolled loop), the 16 x 16 triangle of the chunk stays in registers.  F keeps l (operand of the
// trailing update).  Dynamic shared memory: MW_PANEL_SMEM bytes.
constexpr size_t MW_PANEL_SMEM = sizeof(double) * (2 * MW_NB + MW_NB * MW_NB + 2 * MW_NB + 2 + MW_NB * MW_T);
__global__ void __launch_bounds__(MW_T) mfw_panel_kernel(double* __restrict__ F_all, long long stride, int ld, int shift, int f, int k0, int nb, int j0, int lp0,
                                                         double* __restrict__ Lx_all, size_t nnzL, double* __restrict__ Dv_all, double* __restrict__ Dinv_all, int nk,
                                                         int* __restrict__ fail, const int* __restrict__ active) {
    extern __shared__ __align__(16) double psm[];
    double* col = psm;                            // [2][MW_NB]       double-buffered pivot column (unscaled)
    double* LcT = psm + 2 * MW_NB;                // [MW_NB][MW_NB]   LcT[q * MW_NB + k] = L11(k, q) for k > q, zero elsewhere
    double* dd = LcT + MW_NB * MW_NB;             // [MW_NB]          D
    double* rdd = dd + MW_NB;                     // [MW_NB]          1 / D (fp64 division costs ~30 instructions: one per pivot, not one per entry)
    double* rpiv = rdd + MW_NB;                   // [2]              reciprocal of the current pivot, double-buffered like `col`
    double* wsm = rpiv + 2;                       // [MW_NB][MW_T]    private column of every row-solve thread
    const int b = blockIdx.y;
    if (active && !active[b]) return;
    const int tid = threadIdx.x;
    double* F = F_all + (size_t)b * stride + shift;
    double* Lx = Lx_all + (size_t)b * nnzL;
    const int i = tid & (MW_NB - 1), g = tid >> 6;
    double a[MW_NB / 4];                          // a[m] = A11(i, g + 4 (kk + m)) while the pivots 4 kk .. 4 kk + 3 are eliminated
#pragma unroll
    for (int m = 0; m < MW_NB / 4; m++) { const int c = g + 4 * m; a[m] = (c <= i && i < nb) ? F[(size_t)(k0 + i) + (size_t)(k0 + c) * ld] : 0.0; }
    for (int e = tid; e < MW_NB * MW_NB; e += MW_T) LcT[e] = 0.0;
    __syncthreads();
#pragma unroll 1
    for (int kk = 0; kk < MW_NB / 4; kk++) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int k = 4 * kk + j;
            if (k < nb) {                         // uniform over the CTA
                double* ck = col + (j & 1) * MW_NB;
                if (g == j && i >= k) { ck[i] = a[0]; if (i == k) rpiv[j & 1] = 1.0 / a[0]; }      // column k, unscaled: w_i (and d at i = k)
                __syncthreads();
                const double rd = rpiv[j & 1];
                if (i > k && i < nb) {
                    const double li = ck[i] * rd;                           // l_i = w_i / d
#pragma unroll
                    for (int m = 0; m < MW_NB / 4; m++) { const int c = g + 4 * (kk + m); if (c > k && c <= i) a[m] = __dsub_rn(a[m], __dmul_rn(li, ck[c])); }      // A(i, c) -= l_i w_c, no FMA contraction (sparse/ldlt.hpp:151-158)
                    if (g == j) a[0] = li;
                }
            }
        }
        { const int c = 4 * kk + g; if (c < i) LcT[c * MW_NB + i] = (i < nb) ? a[0] : 0.0; else if (c == i) { const double d = (i < nb) ? a[0] : 1.0; dd[i] = d; rdd[i] = 1.0 / d; } }
#pragma unroll
        for (int m = 0; m + 1 < MW_NB / 4; m++) a[m] = a[m + 1];
        a[MW_NB / 4 - 1] = 0.0;
    }
    __syncthreads();
    if (blockIdx.x == 0) {
        for (int k = tid; k < nb; k += MW_T) {
            const double d = dd[k];
            if (d == 0.0 && fail[b] == 0) fail[b] = j0 + k0 + k + 1;           // ldlt.hpp:161
            Dv_all[(size_t)b * nk + j0 + k0 + k] = d; Dinv_all[(size_t)b * nk + j0 + k0 + k] = 1.0 / d;
        }
        for (int e = tid; e < nb * nb; e += MW_T) {
            const int r = e % nb, c = e / nb;
            if (r > c) Lx[mfw_colbase(lp0, f, k0 + c) + (k0 + r)] = LcT[c * MW_NB + r];
        }
    }
    const int r0 = k0 + nb;
    const int row = r0 + blockIdx.x * MW_T + tid;
    if (row >= f) return;
    constexpr int CH = 16;
#pragma unroll 1
    for (int cb = 0; cb < MW_NB / CH; cb++) {
        if (cb * CH >= nb) break;
        double acc[CH];
#pragma unroll
        for (int j = 0; j < CH; j++) { const int k = cb * CH + j; acc[j] = (k < nb) ? F[(size_t)row + (size_t)(k0 + k) * ld] : 0.0; }
#pragma unroll 2
        for (int q = 0; q < cb * CH; q++) {       // earlier chunks: w_q from the private column, L11(k, q) broadcast as double2
            const double wq = wsm[q * MW_T + tid];
            const double2* lrow = reinterpret_cast<const double2*>(LcT + q * MW_NB + cb * CH);
#pragma unroll
            for (int j = 0; j < CH; j += 2) { const double2 l2 = lrow[j / 2]; acc[j] = __dsub_rn(acc[j], __dmul_rn(wq, l2.x)); acc[j + 1] = __dsub_rn(acc[j + 1], __dmul_rn(wq, l2.y)); }
        }
#pragma unroll
        for (int j = 0; j < CH; j++) {            // the chunk's own triangle
#pragma unroll
            for (int jj = 0; jj < j; jj++) acc[j] = __dsub_rn(acc[j], __dmul_rn(acc[jj], LcT[(cb * CH + jj) * MW_NB + cb * CH + j]));
            wsm[(cb * CH + j) * MW_T + tid] = acc[j];
        }
#pragma unroll
        for (int j = 0; j < CH; j++) {
            const int k = cb * CH + j;
            if (k < nb) {
                const double l = acc[j] * rdd[k];
                F[(size_t)row + (size_t)(k0 + k) * ld] = l;
                Lx[mfw_colbase(lp0, f, k0 + k) + row] = l;
            }
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
// Inverses of the sb x sb unit-lower diagonal blocks of a wide supernode (computed once per factorisation, one CTA per block):
// the blocked solves below apply them as mat-vecs instead of running sb dependent substitution steps per block (the dense
// backend does the same with its 32 x 32 blocks).  Tcm[blk][k * sb + r] = inv(r, k) (column-major), Trm[blk][r * sb + k] = inv(r, k).
// smem: sb * (sb + 1) doubles.
__global__ void __launch_bounds__(MW_T) mfw_block_inverse_kernel(int ws, int f, int lp0, int sb, const double* __restrict__ Lx_all, size_t nnzL,
                                                                 double* __restrict__ Tcm_all, double* __restrict__ Trm_all, long long tinv_stride, long long tinv_off,
                                                                 const int* __restrict__ active) {
    extern __shared__ __align__(16) double sm[];
    const int b = blockIdx.y;
    if (active && !active[b]) return;
    const int tid = threadIdx.x, t = blockIdx.x;
    const int c0 = t * sb, cn = min(sb, ws - c0), ldx = sb + 1;
    const double* Lx = Lx_all + (size_t)b * nnzL;
    double* X = sm;
    for (int e = tid; e < sb * sb; e += MW_T) {
        const int i = e % sb, k = e / sb;
        X[i * ldx + k] = (i < cn && k < cn && i > k) ? Lx[mfw_colbase(lp0, f, c0 + k) + (c0 + i)] : (i == k ? 1.0 : 0.0);
    }
    __syncthreads();
    for (int i = 1; i < cn; i++) {               // row i of the inverse from rows < i:  X(i, j) = - sum_{k=j}^{i-1} T(i, k) X(k, j)
        double acc = 0.0;
        if (tid < i) for (int k = tid; k < i; k++) acc += X[i * ldx + k] * X[k * ldx + tid];
        __syncthreads();
        if (tid < i) X[i * ldx + tid] = -acc;
        __syncthreads();
    }
    double* Tcm = Tcm_all + (size_t)b * tinv_stride + tinv_off + (size_t)t * sb * sb;
    double* Trm = Trm_all + (size_t)b * tinv_stride + tinv_off + (size_t)t * sb * sb;
    for (int e = tid; e < sb * sb; e += MW_T) { const int r = e % sb, k = e / sb; Tcm[e] = X[r * ldx + k]; }
    for (int e = tid; e < sb * sb; e += MW_T) { const int k = e % sb, r = e / sb; Trm[e] = X[r * ldx + k]; }
}
// forward, wide supernode, part 2 (one launch per column block [gc0, gc0 + gcn), gcn = 0 for the first launch): CTA 0 owns the
// rows of the NEXT block [rb, rb + sb), rb = gc0 + gcn; CTAs c >= 1 own 64-row slabs after it.  Every CTA subtracts
// L[rows, block] y[block] (y[block] is final since the previous launch); threads split the block's columns into MW_TS / rows
// groups whose partial sums are added in a fixed order.  CTA 0 then multiplies its rows by the inverse of their diagonal
// block, which makes them final for the next launch.  sb: power of two <= 128.  smem: (MW_TS + sb) doubles.
__global__ void __launch_bounds__(MW_TS) mfw_fwd_block_kernel(int j0, int ws, int f, int lp0, int gc0, int gcn, int sb, const double* __restrict__ Lx_all, size_t nnzL,
                                                             const double* __restrict__ Tcm_all, long long tinv_stride, long long tinv_off,
                                                             double* __restrict__ work_all, int nk, const int* __restrict__ active) {
    extern __shared__ __align__(16) double sm[];
    const int b = blockIdx.y;
    if (active && !active[b]) return;
    const int tid = threadIdx.x;
    const double* Lx = Lx_all + (size_t)b * nnzL;
    double* w = work_all + (size_t)b * nk + j0;
    double* part = sm;               // MW_TS
    double* xs = sm + MW_TS;          // sb
    const int rb = gc0 + gcn, sr = min(64, sb);
    const int rows = blockIdx.x == 0 ? sb : sr;
    const int r0 = blockIdx.x == 0 ? rb : rb + sb + (blockIdx.x - 1) * sr;
    const int rn = min(rows, ws - r0);
    if (rn <= 0) return;
    const int groups = MW_TS / rows, grp = tid / rows, r = tid - grp * rows;
    {
        double acc = 0.0;
        if (r < rn && gcn > 0) {
            const int i = r0 + r;
            const int kh = (gcn + groups - 1) / groups, ka = gc0 + grp * kh, kb = min(gc0 + gcn, ka + kh);
            size_t base = mfw_colbase(lp0, f, ka) + i;
#pragma unroll 8
            for (int k = ka; k < kb; k++) { acc += Lx[base] * w[k]; base += (size_t)(f - k - 2); }
        }
        part[tid] = acc;
    }
    __syncthreads();
    double x = 0.0;
    if (grp == 0 && r < rn) {
        double tot = 0.0;
        for (int g2 = 0; g2 < groups; g2++) tot += part[g2 * rows + r];
        x = w[r0 + r] - tot;
    }
    if (blockIdx.x != 0) {
        if (grp == 0 && r < rn) w[r0 + r] = x;
        return;
    }
    if (grp == 0) xs[r] = r < rn ? x : 0.0;
    __syncthreads();
    {   // y_blk = inv(T_blk) x_blk, lower triangular: columns k <= r
        const double* Tc = Tcm_all + (size_t)b * tinv_stride + tinv_off + (size_t)(r0 / sb) * sb * sb;
        double acc = 0.0;
        if (r < rn) {
            const int kh = (rn + groups - 1) / groups, ka = grp * kh, kb = min(min(rn, r + 1), ka + kh);
#pragma unroll 8
            for (int k = ka; k < kb; k++) acc += Tc[(size_t)k * sb + r] * xs[k];
        }
        part[tid] = acc;
    }
    __syncthreads();
    if (grp == 0 && r < rn) {
        double tot = 0.0;
        for (int g2 = 0; g2 < groups; g2++) tot += part[g2 * rows + r];
        w[r0 + r] = tot;
    }
}
// backward, wide supernode (one launch per column block [c0, c0 + cn), last block first): CTA k computes
// dot_k = sum_{i >= c0+cn} L(i, c0+k) x_i  (rows of the triangle below the block, then the update rows through their row
// indices); the CTA that finishes last subtracts the dots and applies the transposed inverse of the diagonal block.
// smem: (MW_TS + sb + 32) doubles.
__global__ void __launch_bounds__(MW_TS) mfw_bwd_block_kernel(int j0, int ws, int f, int lp0, const int* __restrict__ li_u, int c0, int cn, int sb,
                                                             const double* __restrict__ Lx_all, size_t nnzL, const double* __restrict__ Trm_all, long long tinv_stride,
                                                             long long tinv_off, double* __restrict__ work_all, int nk,
                                                             double* __restrict__ tmp_all, unsigned* __restrict__ counter, const int* __restrict__ active) {
    extern __shared__ __align__(16) double sm[];
    __shared__ int is_last;
    const int b = blockIdx.y;
    if (active && !active[b]) return;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const double* Lx = Lx_all + (size_t)b * nnzL;
    double* wg = work_all + (size_t)b * nk;
    double* tmp = tmp_all + (size_t)b * sb;
    double* part = sm;               // MW_TS
    double* vs = sm + MW_TS;          // sb
    double* red = vs + sb;           // 32
    const int k = blockIdx.x;
    {
        const size_t base = mfw_colbase(lp0, f, c0 + k);
        double acc = 0.0;
#pragma unroll 4
        for (int i = c0 + cn + tid; i < f; i += MW_TS) {
            const double xi = i < ws ? wg[j0 + i] : wg[li_u[i - ws]];
            acc += Lx[base + i] * xi;
        }
        acc = warp_sum(acc);
        if (lane == 0) red[wid] = acc;
        __syncthreads();
        if (tid == 0) {
            double t2 = 0.0;
            for (int w2 = 0; w2 < MW_TS / 32; w2++) t2 += red[w2];
            tmp[k] = t2;
            __threadfence();
            const unsigned prev = atomicAdd(counter + b, 1u);
            is_last = (prev == (unsigned)(cn - 1));
        }
        __syncthreads();
    }
    if (!is_last) return;
    __threadfence();
    if (tid < sb) vs[tid] = tid < cn ? wg[j0 + c0 + tid] - __ldcg(tmp + tid) : 0.0;
    __syncthreads();
    {   // x_blk = inv(T_blk)^T v:  x_k = sum_{i >= k} inv(i, k) v_i ; Trm[i * sb + k] is coalesced over k
        const double* Tr = Trm_all + (size_t)b * tinv_stride + tinv_off + (size_t)(c0 / sb) * sb * sb;
        const int groups = MW_TS / sb, grp = tid / sb, kk = tid - grp * sb;
        double acc = 0.0;
        if (kk < cn) {
            const int ih = (cn + groups - 1) / groups, ia = max(kk, grp * ih), ib = min(cn, (grp + 1) * ih);
#pragma unroll 8
            for (int i = ia; i < ib; i++) acc += Tr[(size_t)i * sb + kk] * vs[i];
        }
        part[tid] = acc;
        __syncthreads();
        if (grp == 0 && kk < cn) {
            double tot = 0.0;
            for (int g2 = 0; g2 < groups; g2++) tot += part[g2 * sb + kk];
            wg[j0 + c0 + kk] = tot;
        }
    }
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
