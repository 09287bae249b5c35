// piqp_b200/csrc/sparse_frontal.cuh -- supernodal multifrontal LDL^T and supernodal triangular solves, one CTA per QP.
//
// Replaces LDLt::factorize_numeric_upper_triangular / solve_inplace (include/piqp/sparse/ldlt.hpp:101-218) for a batch of
// QPs that share one pattern.  The reference's up-looking algorithm is scalar and serial; a batch offers instance
// parallelism (one CTA per instance, no kernel launch per column or level) and each supernode offers dense work:
//
//   mf_factor_kernel : walks the supernodes in postorder.  For supernode s with columns j0..j1 and update rows U the
//                      FRONT (|s|+|U|)^2 is built in shared memory (scatter of the permuted KKT entries through a
//                      precomputed position map, extend-add of the children's update matrices through precomputed
//                      relative indices), its first |s| pivots are eliminated right-looking by the whole CTA
//                      (D_k = F_kk, L_ik = F_ik / D_k, F_ic -= F_ik L_ck), the L columns / D go to HBM in the CSC layout
//                      of L, the Schur complement goes on a per-instance stack for the parent.  Fronts that do not fit in
//                      shared memory use a per-instance HBM/L2 scratch front with the same code.
//   mf_solve_kernel  : permuted rhs in shared memory; forward sweep per supernode = unit-lower solve with the |s| x |s|
//                      triangle, then x_U -= L_Us x_s; D^-1; backward sweep = x_s -= L_Us^T x_U (one warp per column,
//                      shuffle reduction), then the transposed triangle.  Deterministic: no atomics anywhere.
#pragma once
#include "common.cuh"

namespace b200 {

struct MfDev {
    const int *hdr;        // [nsup][8]: j0, ws, us, Lp[j0], asm_begin, asm_cnt, child_begin, child_cnt
    const int *crec;       // [children][4]: us of the child, rel_begin, upd_off (lo, hi)
    const int *rel_idx, *asm_pos, *Li, *perm;
    const long long* upd_off;
    int nsup, nk, n, p, m, front_smem_rows, fmax;
    long long upd_total;
    size_t nnzL, nnzPK;
    int big_right_looking; // 1: fronts beyond shared memory use the right-looking blocked elimination (B200_MF_BIG=right), 0: the left-looking one
    long long* prof;       // optional [8] phase clocks of CTA 0 (B200_MF_PROF=1): zero+scatter, extend-add, eliminate (smem), eliminate (HBM), Schur store
};

constexpr int MF_T = 256;
constexpr int MF_NB = 32;     // panel width of the blocked elimination used for fronts that live in HBM/L2
constexpr int MF_TS = 64;     // trailing-update tile

// ---- elimination of the first ws pivots of a front held in SHARED memory (right-looking, one pivot per step)
__device__ __forceinline__ void mf_eliminate_smem(double* F, int f, int ws, int j0, int lp0, double* lcol, double* Lx, double* Dv, double* Dinv, int* failb) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr int NW = MF_T / 32;
    for (int k = 0; k < ws; k++) {
        const double d = F[k + k * f];
        const double* Fk = F + k * f;
        double* Lcolumn = Lx + (lp0 + k * (f - 1) - (k * (k - 1)) / 2) - (k + 1);      // L(i, j0+k) for front row i > k sits at Lcolumn[i]
        for (int i = k + 1 + tid; i < f; i += MF_T) { const double l = Fk[i] / d; lcol[i] = l; Lcolumn[i] = l; }
        if (tid == 0) {
            if (d == 0.0 && *failb == 0) *failb = j0 + k + 1;           // ldlt.hpp:161
            Dv[j0 + k] = d; Dinv[j0 + k] = 1.0 / d;
        }
        __syncthreads();
        // multiply and subtract are rounded SEPARATELY, like the reference, which forbids FMA contraction in its LDL^T
        // (sparse/ldlt.hpp:151-158).  On the numerically chaotic Maros-Meszaros problems a fused update changes the outcome: the
        // oracle with this very loop restated serially solves QBEACONF / QRECIPE in 17 / 22 iterations without contraction and
        // needs 250 (max_iter) / 146 with it (oracle/experimental_multifrontal.hpp).  The loop is bound by
        // its three shared-memory accesses per element, not by the extra FP64 instruction.
        for (int c = k + 1 + wid; c < f; c += NW) {
            const double lc = lcol[c];
            double* Fc = F + c * f;
#pragma unroll 4
            for (int i = c + lane; i < f; i += 32) Fc[i] = __dsub_rn(Fc[i], __dmul_rn(Fk[i], lc));
        }
        __syncthreads();
    }
}

// ---- blocked elimination of a front held in HBM/L2 (products and sums rounded separately, like mf_eliminate_smem): panels of MF_NB pivots; per panel (1) LDL^T of the nb x nb diagonal
// block in shared memory by one warp, (2) one thread per row below solves its row against the block (w = a L11^-T D, l = w / d),
// (3) the trailing lower triangle is updated tile by tile (64x64, 4x4 per thread) from shared-memory copies of W and L.
// sm: scratch of at least 2 * MF_TS * MF_NB + MF_NB * (MF_NB + 1) + MF_NB doubles.  Wg / Lg: per-instance panel scratch (f x MF_NB each).
__device__ void mf_eliminate_big(double* __restrict__ F, int f, int ws, int j0, int lp0, double* sm, double* __restrict__ Wg, double* __restrict__ Lg,
                                 double* Lx, double* Dv, double* Dinv, int* failb) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    double* A11 = sm;                                   // MF_NB x (MF_NB + 1)
    double* dd = A11 + MF_NB * (MF_NB + 1);             // MF_NB
    double* Wt = dd + MF_NB;                            // MF_NB x MF_TS  (k-major)
    double* Lt = Wt + MF_NB * MF_TS;
    constexpr int LDA = MF_NB + 1;
    for (int k0 = 0; k0 < ws; k0 += MF_NB) {
        const int nb = min(MF_NB, ws - k0), r0 = k0 + nb, R = f - r0;
        // (1) diagonal block
        for (int e = tid; e < nb * nb; e += MF_T) { const int i = e % nb, c = e / nb; A11[i + c * LDA] = (i >= c) ? F[(size_t)(k0 + i) + (size_t)(k0 + c) * f] : 0.0; }
        __syncthreads();
        if (wid == 0) {
            for (int k = 0; k < nb; k++) {
                const double d = A11[k + k * LDA];
                const double wi = (lane > k && lane < nb) ? A11[lane + k * LDA] : 0.0;     // unscaled column k
                const double li = wi / d;
                __syncwarp();
                if (lane > k && lane < nb) A11[lane + k * LDA] = li;
                __syncwarp();
                if (lane > k && lane < nb) for (int c = k + 1; c <= lane; c++) A11[lane + c * LDA] = __dsub_rn(A11[lane + c * LDA], __dmul_rn(wi, A11[c + k * LDA]));   // no FMA contraction (sparse/ldlt.hpp:151-158)
                if (lane == 0) dd[k] = d;
                __syncwarp();
            }
        }
        __syncthreads();
        for (int k = tid; k < nb; k += MF_T) {
            const double d = dd[k];
            if (d == 0.0 && *failb == 0) *failb = j0 + k0 + k + 1;
            Dv[j0 + k0 + k] = d; Dinv[j0 + k0 + k] = 1.0 / d;
        }
        for (int e = tid; e < nb * nb; e += MF_T) {
            const int i = e % nb, c = e / nb;
            if (i > c) { const int kk = k0 + c; Lx[(lp0 + kk * (f - 1) - (kk * (kk - 1)) / 2) - (kk + 1) + (k0 + i)] = A11[i + c * LDA]; }
        }
        // (2) rows below the block
        for (int ii = tid; ii < R; ii += MF_T) {
            const int i = r0 + ii;
            double w[MF_NB];
#pragma unroll
            for (int k = 0; k < MF_NB; k++) w[k] = (k < nb) ? F[(size_t)i + (size_t)(k0 + k) * f] : 0.0;
#pragma unroll
            for (int k = 0; k < MF_NB; k++) {
                if (k < nb) {
                    double acc = w[k];
#pragma unroll
                    for (int q = 0; q < k; q++) acc = __dsub_rn(acc, __dmul_rn(w[q], A11[k + q * LDA]));
                    w[k] = acc;
                }
            }
#pragma unroll
            for (int k = 0; k < MF_NB; k++) {
                if (k < nb) {
                    const double l = w[k] / dd[k];
                    const int kk = k0 + k;
                    Wg[(size_t)k * R + ii] = w[k]; Lg[(size_t)k * R + ii] = l;
                    Lx[(lp0 + kk * (f - 1) - (kk * (kk - 1)) / 2) - (kk + 1) + i] = l;
                }
            }
        }
        __syncthreads();
        // (3) trailing update F[i, c] -= sum_k W[i, k] L[c, k],  r0 <= c <= i < f
        const int nt = (R + MF_TS - 1) / MF_TS;
        const int ti = tid % 16, tc = tid / 16;
        for (int bi = 0; bi < nt; bi++) {
            for (int e = tid; e < nb * MF_TS; e += MF_T) { const int k = e / MF_TS, rr = e % MF_TS, g = bi * MF_TS + rr; Wt[e] = g < R ? Wg[(size_t)k * R + g] : 0.0; }
            for (int bc = 0; bc <= bi; bc++) {
                __syncthreads();
                for (int e = tid; e < nb * MF_TS; e += MF_T) { const int k = e / MF_TS, rr = e % MF_TS, g = bc * MF_TS + rr; Lt[e] = g < R ? Lg[(size_t)k * R + g] : 0.0; }
                __syncthreads();
                double acc[4][4];
#pragma unroll
                for (int a2 = 0; a2 < 4; a2++)
#pragma unroll
                    for (int c2 = 0; c2 < 4; c2++) acc[a2][c2] = 0.0;
                for (int k = 0; k < nb; k++) {
                    const double4 wv = *reinterpret_cast<const double4*>(Wt + k * MF_TS + ti * 4);
                    const double4 lv = *reinterpret_cast<const double4*>(Lt + k * MF_TS + tc * 4);
                    const double wa[4] = {wv.x, wv.y, wv.z, wv.w}, la[4] = {lv.x, lv.y, lv.z, lv.w};
#pragma unroll
                    for (int a2 = 0; a2 < 4; a2++)
#pragma unroll
                        for (int c2 = 0; c2 < 4; c2++) acc[a2][c2] = __dadd_rn(acc[a2][c2], __dmul_rn(wa[a2], la[c2]));
                }
#pragma unroll
                for (int c2 = 0; c2 < 4; c2++) {
                    const int gc = bc * MF_TS + tc * 4 + c2;
#pragma unroll
                    for (int a2 = 0; a2 < 4; a2++) {
                        const int gi = bi * MF_TS + ti * 4 + a2;
                        if (gi < R && gc <= gi) F[(size_t)(r0 + gi) + (size_t)(r0 + gc) * f] -= acc[a2][c2];
                    }
                }
            }
            __syncthreads();
        }
    }
}

// ---- LEFT-LOOKING blocked elimination of a front held in HBM/L2 (default for fronts beyond shared memory).  The assembled front F
// is only READ: for every block of <= 32 columns, one thread per row accumulates  F(row, block) - sum_{k < kmax} L(row, k) D_k L(block, k)
// in 32 registers (L from the CSC storage of L, coalesced over rows; the D L rows of the block staged in shared memory, 64 pivots at
// a time), then either finishes the block's pivots (LDL^T of the 32 x 32 diagonal block by one warp, every row solved against it in
// registers) and writes L / D, or -- for the columns behind the supernode -- writes the Schur complement straight onto the update
// stack.  Compared with the right-looking variant above, the front is never read-modified-written (ws / 32 sweeps over f^2 / 2
// entries saved) and the inner loop is pure register arithmetic: one load per 32 multiply-subtract pairs.  Sums run over k in
// increasing order with products and differences rounded separately: the reference's own order (sparse/ldlt.hpp:139-158).
// sm: >= MFL_KC * 32 + 32 * 33 + 32 doubles.
constexpr int MFL_KC = 64;
__device__ void mf_eliminate_left(const double* __restrict__ F, int f, int ws, int us, int j0, int lp0, double* sm, double* __restrict__ Lx,
                                  double* __restrict__ Dv, double* __restrict__ Dinv, int* failb, double* __restrict__ U) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    double* Wt = sm;                                    // [MFL_KC][32]: Wt[k][j] = D_k L(c0 + j, k)
    double* A11 = Wt + MFL_KC * 32;                     // [32][33]
    double* dd = A11 + 32 * 33;                         // [32]
    constexpr int LDA = 33;
    auto colbase = [&](int k) { return (size_t)((long long)lp0 + (long long)k * (f - 1) - ((long long)k * (k - 1)) / 2 - (k + 1)); };
    int c0 = 0;
    while (c0 < f) {
        const bool panel = c0 < ws;
        const int nb = panel ? min(32, ws - c0) : min(32, f - c0);
        const int kmax = min(c0, ws);
        for (int r0 = c0; r0 < f; r0 += MF_T) {
            const int row = r0 + tid;
            const bool live = row < f;
            double acc[32];
#pragma unroll
            for (int j = 0; j < 32; j++) acc[j] = (live && j < nb && row >= c0 + j) ? F[(size_t)row + (size_t)(c0 + j) * f] : 0.0;
            for (int k0 = 0; k0 < kmax; k0 += MFL_KC) {
                const int kc = min(MFL_KC, kmax - k0);
                __syncthreads();                                   // previous chunk consumed
                for (int e = tid; e < kc * 32; e += MF_T) {
                    const int k = e >> 5, j = e & 31;
                    Wt[e] = j < nb ? Dv[j0 + k0 + k] * Lx[colbase(k0 + k) + (c0 + j)] : 0.0;
                }
                __syncthreads();
                if (live) {
                    // L(row, k) comes from L2 (~0.5 us round trip): eight loads are kept in flight one group ahead of the arithmetic
                    constexpr int PFK = 8;
                    double a[PFK], an[PFK];
#pragma unroll
                    for (int u = 0; u < PFK; u++) a[u] = u < kc ? Lx[colbase(k0 + u) + row] : 0.0;
                    for (int k = 0; k < kc; k += PFK) {
#pragma unroll
                        for (int u = 0; u < PFK; u++) an[u] = (k + PFK + u < kc) ? Lx[colbase(k0 + k + PFK + u) + row] : 0.0;
#pragma unroll
                        for (int u = 0; u < PFK; u++) {
                            if (k + u < kc) {
                                const double2* w2 = reinterpret_cast<const double2*>(Wt + (k + u) * 32);
#pragma unroll
                                for (int j = 0; j < 16; j++) {
                                    const double2 w = w2[j];
                                    acc[2 * j] = __dsub_rn(acc[2 * j], __dmul_rn(a[u], w.x));
                                    acc[2 * j + 1] = __dsub_rn(acc[2 * j + 1], __dmul_rn(a[u], w.y));
                                }
                            }
                        }
#pragma unroll
                        for (int u = 0; u < PFK; u++) a[u] = an[u];
                    }
                }
            }
            if (!panel) {                                          // Schur complement block -> update stack (lower part)
                if (live) {
#pragma unroll
                    for (int j = 0; j < 32; j++) if (j < nb && row >= c0 + j) U[(size_t)(row - ws) + (size_t)(c0 + j - ws) * us] = acc[j];
                }
                continue;
            }
            if (r0 == c0) {                                        // first row tile: holds the rows of the diagonal block
                __syncthreads();
                if (tid < nb) {
#pragma unroll
                    for (int j = 0; j < 32; j++) if (j < nb) A11[tid + j * LDA] = (j <= tid) ? acc[j] : 0.0;
                }
                __syncthreads();
                if (wid == 0) {                                    // LDL^T of the nb x nb block, right-looking, one warp (same arithmetic as mf_eliminate_big)
                    for (int k = 0; k < nb; k++) {
                        const double d = A11[k + k * LDA];
                        const double wi = (lane > k && lane < nb) ? A11[lane + k * LDA] : 0.0;
                        const double li = wi / d;
                        __syncwarp();
                        if (lane > k && lane < nb) A11[lane + k * LDA] = li;
                        __syncwarp();
                        if (lane > k && lane < nb) for (int c = k + 1; c <= lane; c++) A11[lane + c * LDA] = __dsub_rn(A11[lane + c * LDA], __dmul_rn(wi, A11[c + k * LDA]));
                        if (lane == 0) dd[k] = d;
                        __syncwarp();
                    }
                }
                __syncthreads();
                for (int k = tid; k < nb; k += MF_T) {
                    const double d = dd[k];
                    if (d == 0.0 && *failb == 0) *failb = j0 + c0 + k + 1;       // ldlt.hpp:161
                    Dv[j0 + c0 + k] = d; Dinv[j0 + c0 + k] = 1.0 / d;
                }
                for (int e = tid; e < nb * nb; e += MF_T) {
                    const int i = e % nb, c = e / nb;
                    if (i > c) Lx[colbase(c0 + c) + (c0 + i)] = A11[i + c * LDA];
                }
            }
            if (live && row >= c0 + nb) {                          // rows below the block: w = a L11^-T, l = w / d
#pragma unroll
                for (int k = 0; k < 32; k++) {
                    if (k < nb) {
                        double a = acc[k];
#pragma unroll
                        for (int q = 0; q < k; q++) a = __dsub_rn(a, __dmul_rn(acc[q], A11[k + q * LDA]));
                        acc[k] = a;
                    }
                }
#pragma unroll
                for (int k = 0; k < 32; k++) if (k < nb) Lx[colbase(c0 + k) + row] = acc[k] / dd[k];
            }
        }
        __syncthreads();                                           // L / D of this block visible to the next block's Wt staging (same CTA: global writes + barrier)
        c0 += nb;
    }
}

// PKasm: the permuted KKT values in ASSEMBLY order (grouped by supernode, see LdltSymbolic::asm_*), so that a front's
// original entries are one contiguous, coalesced read.
__global__ void __launch_bounds__(MF_T) mf_factor_kernel(MfDev M, const double* __restrict__ PKasm_all, double* __restrict__ Lx_all, double* __restrict__ Dv_all,
                                                         double* __restrict__ Dinv_all, double* __restrict__ upd_all, double* __restrict__ big_all,
                                                         double* __restrict__ panel_all, int* __restrict__ fail, const int* __restrict__ active) {
    extern __shared__ __align__(16) double mf_sm[];
    const int b = blockIdx.x;
    if (active && !active[b]) return;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr int NW = MF_T / 32;
    const double* PK = PKasm_all + (size_t)b * M.nnzPK;
    double* Lx = Lx_all + (size_t)b * M.nnzL;
    double* Dv = Dv_all + (size_t)b * M.nk;
    double* Dinv = Dinv_all + (size_t)b * M.nk;
    double* upd = upd_all + (size_t)b * (size_t)M.upd_total;
    double* big = big_all ? big_all + (size_t)b * (size_t)M.fmax * M.fmax : nullptr;
    double* Wg = panel_all ? panel_all + (size_t)b * 2 * (size_t)M.fmax * MF_NB : nullptr;
    double* Lg = Wg ? Wg + (size_t)M.fmax * MF_NB : nullptr;
    const int fpad = (M.fmax + 1) & ~1;
    double* lcol = mf_sm;                                  // fmax doubles
    int* relbuf = reinterpret_cast<int*>(mf_sm + fpad);    // 2 x fmax ints (double buffer of the extend-add)
    double* Fs = mf_sm + ((2 * fpad + 3) & ~3);
    if (tid == 0) fail[b] = 0;
    const int4* hdr4 = reinterpret_cast<const int4*>(M.hdr);
    int4 nh0 = hdr4[0], nh1 = hdr4[1];
    long long noff = M.upd_off[0];

    const bool prof = M.prof != nullptr && b == 0 && tid == 0;
    long long pc[6] = {0, 0, 0, 0, 0, 0}, t_prev = prof ? clock64() : 0;
#define MF_LAP(k) do { if (prof) { const long long t_ = clock64(); pc[k] += t_ - t_prev; t_prev = t_; } } while (0)
    for (int s = 0; s < M.nsup; s++) {
        const int4 h0 = nh0, h1 = nh1;
        const long long my_off = noff;
        if (s + 1 < M.nsup) { nh0 = hdr4[2 * (s + 1)]; nh1 = hdr4[2 * (s + 1) + 1]; noff = M.upd_off[s + 1]; }     // prefetch the next header
        const int j0 = h0.x, ws = h0.y, us = h0.z, lp0 = h0.w, ab = h1.x, an = h1.y, cb = h1.z, cn = h1.w;
        const int f = ws + us;
        const bool in_smem = f <= M.front_smem_rows;
        double* F = in_smem ? Fs : big;
        // original entries of this front: issue the loads before zeroing
        double av[2]; int ap[2];
#pragma unroll
        for (int u = 0; u < 2; u++) { const int t = tid + u * MF_T; if (t < an) { ap[u] = M.asm_pos[ab + t]; av[u] = PK[ab + t]; } }
        if (in_smem) { for (int e = tid; e < f * f; e += MF_T) Fs[e] = 0.0; }
        else { for (size_t e = tid; e < (size_t)f * f; e += MF_T) big[e] = 0.0; }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < 2; u++) if (tid + u * MF_T < an) F[ap[u]] = av[u];
        for (int t = tid + 2 * MF_T; t < an; t += MF_T) F[M.asm_pos[ab + t]] = PK[ab + t];
        MF_LAP(0);
        // extend-add, one child at a time in the fixed (postorder) order: deterministic, no atomics.  The next child's record and
        // relative indices are fetched while the current child's update matrix is added (double-buffered relbuf, one barrier per
        // child); only the lower triangle of an update matrix is read, one warp per column.
        if (cn > 0) {
            int4 crn = reinterpret_cast<const int4*>(M.crec)[cb];
            for (int t = tid; t < crn.x; t += MF_T) relbuf[t] = M.rel_idx[crn.y + t];
            for (int c = 0; c < cn; c++) {
                const int4 cr = crn;
                int* rel = relbuf + (c & 1) * M.fmax;
                int* reln = relbuf + ((c + 1) & 1) * M.fmax;
                if (c + 1 < cn) crn = reinterpret_cast<const int4*>(M.crec)[cb + c + 1];
                __syncthreads();                              // rel ready; the previous child's additions (and the scatter of the original entries) are done
                if (c + 1 < cn) for (int t = tid; t < crn.x; t += MF_T) reln[t] = M.rel_idx[crn.y + t];
                const int uc = cr.x;
                const double* U = upd + (((long long)cr.w << 32) | (unsigned)cr.z);
                for (int col = wid; col < uc; col += NW) {
                    const size_t cbase = (size_t)rel[col] * f;
                    const double* Uc = U + (size_t)col * uc;
                    for (int a = col + lane; a < uc; a += 32) { const size_t pos = (size_t)rel[a] + cbase; if (in_smem) Fs[pos] += Uc[a]; else big[pos] += Uc[a]; }
                }
            }
            __syncthreads();
        }
        if (cn == 0) __syncthreads();
        MF_LAP(1);
        // ---- eliminate the ws pivots of the supernode
        if (in_smem) { mf_eliminate_smem(Fs, f, ws, j0, lp0, lcol, Lx, Dv, Dinv, fail + b); MF_LAP(2); }
        else if (M.big_right_looking) { mf_eliminate_big(big, f, ws, j0, lp0, Fs, Wg, Lg, Lx, Dv, Dinv, fail + b); __syncthreads(); MF_LAP(3); }
        else { mf_eliminate_left(big, f, ws, us, j0, lp0, Fs, Lx, Dv, Dinv, fail + b, upd + my_off); __syncthreads(); MF_LAP(3); }
        // ---- Schur complement -> stack (full us x us square, ld = us; only the lower part is meaningful)
        if (us > 0 && (in_smem || M.big_right_looking)) {       // the left-looking elimination wrote its Schur complement itself
            double* U = upd + my_off;
            for (int col = wid; col < us; col += NW) {
                const double* Fc = F + (size_t)(ws + col) * f + ws;
                for (int a = col + lane; a < us; a += 32) U[a + (size_t)col * us] = Fc[a];
            }
        }
        __syncthreads();
        MF_LAP(4);
    }
#undef MF_LAP
    if (prof) for (int k = 0; k < 5; k++) M.prof[k] += pc[k];
}

// x: permuted work vector (shared memory when it fits, else the HBM work buffer)
__global__ void __launch_bounds__(MF_T) mf_solve_kernel(MfDev M, int x_in_smem, const double* __restrict__ Lx_all, const double* __restrict__ Dinv_all,
                                                        const double* __restrict__ rx, const double* __restrict__ ry, const double* __restrict__ rz,
                                                        double* __restrict__ lx, double* __restrict__ ly, double* __restrict__ lz, double* __restrict__ work_all,
                                                        const int* __restrict__ active) {
    extern __shared__ __align__(16) double mf_sm[];
    const int b = blockIdx.x;
    if (active && !active[b]) return;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr int NW = MF_T / 32;
    const int nk = M.nk, n = M.n, p = M.p, m = M.m;
    const double* Lx = Lx_all + (size_t)b * M.nnzL;
    const double* Dinv = Dinv_all + (size_t)b * nk;
    double* x = x_in_smem ? mf_sm : work_all + (size_t)b * nk;
    __shared__ double red_fallback[NW];
    for (int j = tid; j < nk; j += MF_T) {                  // ordering.hpp:102-124, sparse/kkt.hpp:113-147
        const int v = M.perm[j];
        x[j] = v < n ? rx[(size_t)b * n + v] : (v < n + p ? ry[(size_t)b * p + (v - n)] : rz[(size_t)b * m + (v - n - p)]);
    }
    __syncthreads();
    const int4* hdr4 = reinterpret_cast<const int4*>(M.hdr);
    // column j0+k of a supernode starts at lp0 + k (f-1) - k (k-1) / 2 in L; its last |U| entries are the update rows
    // ---- forward: L y = b.  The first MF_T entries of the NEXT supernode's update panel are prefetched into registers.
    int4 nh = hdr4[0];
    double pl = 0.0; int pu = 0;
    { const int ws = nh.y, us = nh.z; if (ws == 1 && tid < us) { pl = Lx[nh.w + tid]; pu = M.Li[nh.w + tid]; } }
    for (int s = 0; s < M.nsup; s++) {
        const int4 h = nh;
        const double cl = pl; const int cu = pu;
        if (s + 1 < M.nsup) {
            nh = hdr4[2 * (s + 1)];
            if (nh.y == 1 && tid < nh.z) { pl = Lx[nh.w + tid]; pu = M.Li[nh.w + tid]; }
        }
        const int j0 = h.x, ws = h.y, us = h.z, lp0 = h.w, f = ws + us;
        if (ws == 1) {
            if (us > 0) {
                const double xk = x[j0];
                if (tid < us) x[cu] -= cl * xk;
                for (int t = tid + MF_T; t < us; t += MF_T) x[M.Li[lp0 + t]] -= Lx[lp0 + t] * xk;
                __syncthreads();
            }
            continue;
        }
        for (int k = 0; k + 1 < ws; k++) {
            const double xk = x[j0 + k];
            const double* Lc = Lx + (lp0 + k * (f - 1) - (k * (k - 1)) / 2);            // rows j0+k+1 .. j1 first
            for (int i = k + 1 + tid; i < ws; i += MF_T) x[j0 + i] -= Lc[i - k - 1] * xk;
            __syncthreads();
        }
        if (us > 0) {
            const int* U = M.Li + (lp0 + (ws - 1) * (f - 1) - ((ws - 1) * (ws - 2)) / 2);
            for (int t = tid; t < us; t += MF_T) {
                double acc = 0.0;
                for (int k = 0; k < ws; k++) acc += Lx[(lp0 + k * (f - 1) - (k * (k - 1)) / 2) + (ws - 1 - k) + t] * x[j0 + k];
                x[U[t]] -= acc;
            }
            __syncthreads();
        }
    }
    for (int j = tid; j < nk; j += MF_T) x[j] *= Dinv[j];
    __syncthreads();
    // ---- backward: L^T x = y  (same register prefetch, walking the supernodes in reverse)
    nh = hdr4[2 * (M.nsup - 1)];
    pl = 0.0; pu = 0;
    if (nh.y == 1 && tid < nh.z) { pl = Lx[nh.w + tid]; pu = M.Li[nh.w + tid]; }
    for (int s = M.nsup - 1; s >= 0; s--) {
        const int4 h = nh;
        const double cl = pl; const int cu = pu;
        if (s > 0) {
            nh = hdr4[2 * (s - 1)];
            if (nh.y == 1 && tid < nh.z) { pl = Lx[nh.w + tid]; pu = M.Li[nh.w + tid]; }
        }
        const int j0 = h.x, ws = h.y, us = h.z, lp0 = h.w, f = ws + us;
        if (ws == 1) {
            if (us > 0) {
                double acc = tid < us ? cl * x[cu] : 0.0;
                for (int t = tid + MF_T; t < us; t += MF_T) acc += Lx[lp0 + t] * x[M.Li[lp0 + t]];
                // fixed-tree block reduction (deterministic)
                acc = warp_sum(acc);
                double* red = red_fallback;
                if (lane == 0) red[wid] = acc;
                __syncthreads();
                if (tid == 0) { double t2 = 0.0; for (int w2 = 0; w2 < (us + 31) / 32 && w2 < NW; w2++) t2 += red[w2]; x[j0] -= t2; }
                __syncthreads();
            }
            continue;
        }
        if (us > 0) {
            const int* U = M.Li + (lp0 + (ws - 1) * (f - 1) - ((ws - 1) * (ws - 2)) / 2);
            for (int k = wid; k < ws; k += NW) {
                const double* Lc = Lx + (lp0 + k * (f - 1) - (k * (k - 1)) / 2) + (ws - 1 - k);
                double acc = 0.0;
                for (int t = lane; t < us; t += 32) acc += Lc[t] * x[U[t]];
                acc = warp_sum(acc);
                if (lane == 0) x[j0 + k] -= acc;
            }
            __syncthreads();
        }
        for (int k = ws - 2; k >= 0; k--) {
            if (wid == 0) {
                const double* Lc = Lx + (lp0 + k * (f - 1) - (k * (k - 1)) / 2);
                double acc = 0.0;
                for (int i = k + 1 + lane; i < ws; i += 32) acc += Lc[i - k - 1] * x[j0 + i];
                acc = warp_sum(acc);
                if (lane == 0) x[j0 + k] -= acc;
            }
            __syncthreads();
        }
    }
    for (int j = tid; j < nk; j += MF_T) {
        const int v = M.perm[j];
        const double val = x[j];
        if (v < n) lx[(size_t)b * n + v] = val; else if (v < n + p) ly[(size_t)b * p + (v - n)] = val; else lz[(size_t)b * m + (v - n - p)] = val;
    }
}


// ------------------------------------------------------------------------------------------------------------------
// mf_solve_ring_kernel: the same supernodal sweeps with the L panels STREAMED through shared memory.  In mf_solve_kernel every
// triangle step of a multi-column supernode waits for an L2 round trip; but L does not depend on x, so the panel of column block
// s+1 (one contiguous range of Lx, because the columns of a supernode are adjacent in L's CSC storage) and its row indices are
// copied with cp.async into the other half of a double buffer while block s is processed.  Supernodes are cut into column blocks
// whose panel fits the buffer (host: SparseLdltBatchedKKT builds `shdr`).
// shdr[blk] = {j0, wsb, nbelow, lx_off, len, li_off, 0, 0}
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mf_cp8(double* dst, const double* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void mf_cp4(int* dst, const int* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__global__ void __launch_bounds__(MF_T) mf_solve_ring_kernel(MfDev M, const int* __restrict__ shdr, int nblk, int PB, int RB, const double* __restrict__ Lx_all,
                                                             const double* __restrict__ Dinv_all, const double* __restrict__ rx, const double* __restrict__ ry,
                                                             const double* __restrict__ rz, double* __restrict__ lx, double* __restrict__ ly, double* __restrict__ lz,
                                                             const int* __restrict__ active) {
    extern __shared__ __align__(16) double mf_sm[];
    __shared__ double red[MF_T / 32];
    const int b = blockIdx.x;
    if (active && !active[b]) return;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr int NW = MF_T / 32;
    const int nk = M.nk, n = M.n, p = M.p, m = M.m;
    const double* Lx = Lx_all + (size_t)b * M.nnzL;
    const double* Dinv = Dinv_all + (size_t)b * nk;
    double* x = mf_sm;
    double* pan = mf_sm + ((nk + 1) & ~1);
    int* ridx = reinterpret_cast<int*>(pan + 2 * (size_t)PB);
    const int4* hdr4 = reinterpret_cast<const int4*>(shdr);
    // headers are prefetched in registers two blocks ahead, panels one block ahead
    auto prefetch = [&](const int4& h0, const int4& h1, int buf) {
        const double* src = Lx + h0.w;
        double* dst = pan + (size_t)buf * PB;
        for (int e = tid; e < h1.x; e += MF_T) mf_cp8(dst + e, src + e);
        const int* isrc = M.Li + h1.y;
        int* idst = ridx + (size_t)buf * RB;
        for (int e = tid; e < h0.z; e += MF_T) mf_cp4(idst + e, isrc + e);
    };
    const int4 zero4 = make_int4(0, 0, 0, 0);
    int4 c0 = nblk > 0 ? hdr4[0] : zero4, c1 = nblk > 0 ? hdr4[1] : zero4;
    int4 n0 = nblk > 1 ? hdr4[2] : zero4, n1 = nblk > 1 ? hdr4[3] : zero4;
    if (nblk > 0) prefetch(c0, c1, 0);
    asm volatile("cp.async.commit_group;\n" ::: "memory");
    for (int j = tid; j < nk; j += MF_T) {
        const int v = M.perm[j];
        x[j] = v < n ? rx[(size_t)b * n + v] : (v < n + p ? ry[(size_t)b * p + (v - n)] : rz[(size_t)b * m + (v - n - p)]);
    }
    // ---- forward
    for (int s = 0; s < nblk; s++) {
        int4 m0 = zero4, m1 = zero4;
        if (s + 2 < nblk) { m0 = hdr4[2 * (s + 2)]; m1 = hdr4[2 * (s + 2) + 1]; }
        if (s + 1 < nblk) prefetch(n0, n1, (s + 1) & 1);
        asm volatile("cp.async.commit_group;\n" ::: "memory");
        asm volatile("cp.async.wait_group 1;\n" ::: "memory");
        __syncthreads();
        const int j0 = c0.x, wsb = c0.y, nb = c0.z, fb = wsb + nb;
        const double* P = pan + (size_t)(s & 1) * PB;
        const int* R = ridx + (size_t)(s & 1) * RB;
        for (int k = 0; k + 1 < wsb; k++) {
            const double xk = x[j0 + k];
            const double* col = P + (k * (fb - 1) - (k * (k - 1)) / 2);
            for (int i = k + 1 + tid; i < wsb; i += MF_T) x[j0 + i] -= col[i - k - 1] * xk;
            __syncthreads();
        }
        for (int t = tid; t < nb; t += MF_T) {
            double acc = 0.0;
            for (int k = 0; k < wsb; k++) acc += P[(k * (fb - 1) - (k * (k - 1)) / 2) + (wsb - 1 - k) + t] * x[j0 + k];
            x[R[t]] -= acc;
        }
        __syncthreads();
        c0 = n0; c1 = n1; n0 = m0; n1 = m1;
    }
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    for (int j = tid; j < nk; j += MF_T) x[j] *= Dinv[j];
    __syncthreads();
    // ---- backward (blocks in reverse; buffer parity follows the block index again)
    c0 = nblk > 0 ? hdr4[2 * (nblk - 1)] : zero4; c1 = nblk > 0 ? hdr4[2 * (nblk - 1) + 1] : zero4;
    n0 = nblk > 1 ? hdr4[2 * (nblk - 2)] : zero4; n1 = nblk > 1 ? hdr4[2 * (nblk - 2) + 1] : zero4;
    if (nblk > 0) prefetch(c0, c1, (nblk - 1) & 1);
    asm volatile("cp.async.commit_group;\n" ::: "memory");
    for (int s = nblk - 1; s >= 0; s--) {
        int4 m0 = zero4, m1 = zero4;
        if (s >= 2) { m0 = hdr4[2 * (s - 2)]; m1 = hdr4[2 * (s - 2) + 1]; }
        if (s > 0) prefetch(n0, n1, (s - 1) & 1);
        asm volatile("cp.async.commit_group;\n" ::: "memory");
        asm volatile("cp.async.wait_group 1;\n" ::: "memory");
        __syncthreads();
        const int j0 = c0.x, wsb = c0.y, nb = c0.z, fb = wsb + nb;
        const double* P = pan + (size_t)(s & 1) * PB;
        const int* R = ridx + (size_t)(s & 1) * RB;
        if (nb > 0) {
            if (wsb == 1 && nb > 64) {                     // one long column: all warps, fixed-tree block reduction
                double acc = 0.0;
                for (int t = tid; t < nb; t += MF_T) acc += P[t] * x[R[t]];
                acc = warp_sum(acc);
                if (lane == 0) red[wid] = acc;
                __syncthreads();
                if (tid == 0) { double t2 = 0.0; for (int w2 = 0; w2 < NW; w2++) t2 += red[w2]; x[j0] -= t2; }
            } else {
                for (int k = wid; k < wsb; k += NW) {
                    const double* col = P + (k * (fb - 1) - (k * (k - 1)) / 2) + (wsb - 1 - k);
                    double acc = 0.0;
                    for (int t = lane; t < nb; t += 32) acc += col[t] * x[R[t]];
                    acc = warp_sum(acc);
                    if (lane == 0) x[j0 + k] -= acc;
                }
            }
            __syncthreads();
        }
        for (int k = wsb - 2; k >= 0; k--) {
            if (wid == 0) {
                const double* col = P + (k * (fb - 1) - (k * (k - 1)) / 2);
                double acc = 0.0;
                for (int i = k + 1 + lane; i < wsb; i += 32) acc += col[i - k - 1] * x[j0 + i];
                acc = warp_sum(acc);
                if (lane == 0) x[j0 + k] -= acc;
            }
            __syncthreads();
        }
        c0 = n0; c1 = n1; n0 = m0; n1 = m1;
    }
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    __syncthreads();
    for (int j = tid; j < nk; j += MF_T) {
        const int v = M.perm[j];
        const double val = x[j];
        if (v < n) lx[(size_t)b * n + v] = val; else if (v < n + p) ly[(size_t)b * p + (v - n)] = val; else lz[(size_t)b * m + (v - n - p)] = val;
    }
}

}  // namespace b200
