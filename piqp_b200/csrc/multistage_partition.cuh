// piqp_b200/csrc/multistage_partition.cuh -- PARALLEL-IN-HORIZON factorisation and solves of the multistage backend.
//
// factor_kkt / solve_llt_in_place (include/piqp/sparse/multistage_kkt.hpp:1253-1352, 1709-1816) walk the N stages of the
// block-tridiagonal KKT one after the other, and so do the warp-chain kernels of multistage_chain.cuh: their time is N x (latency
// of one stage).  Here the horizon is cut at K-1 SEPARATOR stages g_1 < ... < g_{K-1} into K runs of consecutive stages:
//
//      [ run 0 ] g_1 [ run 1 ] g_2 ... g_{K-1} [ run K-1 ]
//
// Without the separators the runs are independent block-tridiagonal chains, so (a partitioned / substructuring Cholesky, i.e. the
// same factorisation under the elimination order  run 0, ..., run K-1, g_1, ..., g_{K-1}):
//   1. msw_factor_chain_kernel in SEGMENT MODE factorises all runs of all QPs at once (grid = batch x K);
//   2. msp_spike_kernel computes the spike Y_s = L_s^-1 K[run s, g_s] of every run s >= 1 (the separator on its LEFT couples only
//      to the first stage of the run, but the forward substitution fills the whole run): d(g_s) right-hand sides per run, one
//      warp each, the run's forward packets staged once per CTA;
//   3. the same kernel accumulates Y^T Y and emits the reduced system on the separators, which is again block tridiagonal with the stage
//      shape of the original chain:  D~_k = D(g_k) + carry(run k-1) - Y_k^T Y_k ,  B~_k = -(B L^-T)(last stage of run k) Y_k[last];
//   4. msw_factor_chain_kernel factorises the reduced chain (K-1 stages).
// Chain length N/K + K instead of N.  The solves follow the same split: msp_fwd_kernel (runs in parallel, also accumulates
// Y_s^T y_s), msp_gather_kernel (reduced right-hand side), msw_solve_kernel on the reduced chain, msp_bwd_kernel (runs in
// parallel: x_s = L_s^-T (y_s - Y_s x(g_s)) with the coupling to g_{s+1} through the ordinary backward packet).
// The spikes ride in the solve packets of the run's stages: forward packet + Y^T block [Dsep x D], backward packet + Y block [D x Dsep].
#pragma once
#include "multistage_chain.cuh"

namespace b200 {

struct MsPart {
    const int* seg_bounds;     // [2K]  stages [i0, i1) of run s
    const int* sep;            // [K-1] separator stages g_1 .. g_{K-1}
    const int* rstart;         // [K]   start of separator k in the reduced vector (last entry = its length)
    const int* roffD;          // [K-1] offsets of D~_k / B~_k in the reduced block storage
    const int* roffB;
    int K, rn, rtotal;
};

__device__ __forceinline__ int msp_yF(const int* m_cls, int i) { const int D = m_cls[i]; return D * D + D * (i > 0 ? m_cls[i - 1] : 0); }
__device__ __forceinline__ int msp_yB(const int* m_cls, int i, int N) { const int D = m_cls[i]; return D * D + D * ((i + 2 < N) ? m_cls[i + 1] : 0); }

constexpr int MSP_R = 4;       // ring depth of the CTA-wide packet ring of the spike kernel
constexpr int MSP_PF = 6;      // packets in flight per warp of msp_fwd / msp_bwd (smem per CTA decides how many runs are resident per SM)

// ---- 2 + 3. spikes and the reduced system: CTA per (QP, run s >= 1), warp j = column j of the separator block on the run's left.
// Warp j substitutes column j of K[run, g] through the run (Y = L^-1 K[run, g]) with the run's forward packets staged once per CTA in
// a shared-memory ring, writes Y / Y^T into the solve packets, and accumulates row j of the Gram matrix Y^T Y from the columns the
// other warps hold in shared memory.  At the end the CTA emits the reduced blocks of separator k = run - 1:
//     D~_k = D(g_k) + carry(run k-1... the run on the left) - Y^T Y ,   B~_k = -(B L^-T)(last stage of this run) Y[last stage]
// smem: ring[MSP_R][slot] | per warp y[3][32] + tmp[32]
__global__ void __launch_bounds__(1024) msp_spike_kernel(MsDev s, MsPart P, int slot_doubles, const double* __restrict__ fac_all,
                                                         double* __restrict__ pk_all, size_t pk_stride, const double* __restrict__ carry_all,
                                                         double* __restrict__ rfac_all, const int* __restrict__ active) {
    extern __shared__ __align__(16) double sp_sm[];
    const int b = blockIdx.x;
    if (active && !active[b]) return;
    const int run = blockIdx.y + 1, k = run - 1;
    const int tid = threadIdx.x, lane = tid & 31, j = tid >> 5, nth = blockDim.x;
    const int N = s.N;
    const int i0 = P.seg_bounds[2 * run], i1 = P.seg_bounds[2 * run + 1], g = i0 - 1;
    double* ring = sp_sm;
    double* ybase = ring + (size_t)MSP_R * slot_doubles;                    // warp w: ybase + 128 w : y[3][32] | tmp[32]
    double* y = ybase + (size_t)j * 128;
    double* tmp = y + 96;
    const double* fac = fac_all + (size_t)b * s.total;
    double* pk = pk_all + (size_t)b * pk_stride;
    const int dg = s.diag[g], og = s.off[g], Dsep = s.cls[g];
    auto issue = [&](int i) {                        // inv(L_i) | B_{i-1} of the forward packet
        const int D = s.cls[i];
        const int sz = D * D + D * (i > 0 ? s.cls[i - 1] : 0);
        const double* src = pk + s.pkF[i];
        double* dst = ring + (size_t)((i - i0) % MSP_R) * slot_doubles;
        for (int e = 2 * tid; e < sz; e += 2 * nth) msw_cp_async16(dst + e, src + e);
    };
    for (int q = 0; q < MSP_R - 1; q++) { if (i0 + q < i1) issue(i0 + q); msw_cp_commit(); }
    // right-hand side: column j of B(g), rows = the first og variables of the run's first stage
    y[lane] = (j < dg && lane < og) ? __ldg(fac + s.offB[g] + lane + (size_t)j * og) : 0.0;
    y[32 + lane] = 0.0; y[64 + lane] = 0.0;
    double gram = 0.0;                               // (Y^T Y)(j, lane)
    auto gram_add = [&](int buf) {                   // += sum_r Y_stage(r, j) Y_stage(r, lane), columns of the other warps from shared memory
        if (lane < Dsep) {
            const double* mine = y + buf * 32;
            const double* other = ybase + (size_t)lane * 128 + buf * 32;
            double a = 0.0;
#pragma unroll 8
            for (int r = 0; r < 32; r++) a += mine[r] * other[r];      // rows >= d of a stage vector are zero
            gram += a;
        }
    };
    for (int i = i0; i < i1; i++) {
        msw_cp_wait<MSP_R - 2>();
        __syncthreads();                             // packet i landed for every thread; every warp finished stage i - 1
        if (i + MSP_R - 1 < i1) issue(i + MSP_R - 1);
        msw_cp_commit();
        const int t = i - i0, cur = (t % 3) * 32, prev = ((t + 2) % 3) * 32, nxt = ((t + 1) % 3) * 32;
        if (t > 0) gram_add((t + 2) % 3);
        const int d = s.diag[i], D = s.cls[i];
        const int PD = t > 0 ? s.cls[i - 1] : 0;
        const double* pkt = ring + (size_t)(t % MSP_R) * slot_doubles;
        if (D == 16) msw_fwd_stage<16>(pkt, PD, y, cur, prev, d, 0, tmp, nullptr, lane);
        else if (D == 8) msw_fwd_stage<8>(pkt, PD, y, cur, prev, d, 0, tmp, nullptr, lane);
        else msw_fwd_stage<32>(pkt, PD, y, cur, prev, d, 0, tmp, nullptr, lane);
        if (lane >= d) y[cur + lane] = 0.0;          // keep the vector zero-padded (the Gram sums run over 32 rows)
        // Y into the packets of this stage, both orientations, zero-padded to the class sizes
        if (j < Dsep && lane < D) {
            const double v = (lane < d && j < dg) ? y[cur + lane] : 0.0;
            pk[s.pkB[i] + msp_yB(s.cls, i, N) + lane + j * D] = v;            // Y  [D x Dsep]
            pk[s.pkF[i] + msp_yF(s.cls, i) + j + lane * Dsep] = v;            // Y^T [Dsep x D]
        }
        // the buffer of stage t - 2 becomes the vector of stage t + 1 (zero right-hand side): every warp read it for its Gram row
        // before the barrier at the top of this iteration
        y[nxt + lane] = 0.0;
        __syncwarp();
    }
    msw_cp_wait<0>();
    __syncthreads();
    const int tl = i1 - 1 - i0, lastbuf = tl % 3;
    gram_add(lastbuf);
    // ---- reduced blocks of separator k (lower triangle of D~, upper part zero like the assembled blocks of the chain)
    double* rfac = rfac_all + (size_t)b * P.rtotal;
    if (j < dg) {
        const int ol = s.off[g - 1];
        const double* carry = carry_all + ((size_t)b * P.K + k) * 1024;      // Schur complement the run on the LEFT left on its coupling rows
        if (lane < dg) {
            double v = 0.0;
            if (j >= lane) {                         // row j, column lane
                v = fac[s.offD[g] + j + (size_t)lane * dg];
                if (j < ol) v += carry[j + 32 * lane];
                v -= gram;
            }
            rfac[P.roffD[k] + j + lane * dg] = v;
        }
        if (k + 1 < P.K - 1) {                       // B~_k(:, j) = -(B L^-T)(last stage) Y_last(:, j)
            const int il = i1 - 1, o2 = s.off[il], D = s.cls[il], dl = s.diag[il];
            const double* BT = pk + s.pkB[il] + D * D;                        // B^T [D x ND]: BT[q + r2 * D], written by the chain kernel
            const double* yl = y + lastbuf * 32;
            if (lane < o2) {
                double a = 0.0;
                for (int q = 0; q < dl; q++) a += BT[q + lane * D] * yl[q];
                rfac[P.roffB[k] + lane + j * o2] = -a;
            }
        }
    }
}

// ---- solves.  smem of msp_fwd / msp_bwd: xs[seg_len_max + 96] | tmp[32] | z[32] | ring[MSP_PF][slot]; the meta block is read from global memory
__global__ void __launch_bounds__(32) msp_fwd_kernel(MsDev s, MsPart P, int slot_doubles, int seg_len_max, const double* __restrict__ pk_all, size_t pk_stride,
                                                     double* __restrict__ X, double* __restrict__ zbuf, const int* __restrict__ active) {
    extern __shared__ __align__(16) double xs[];
    const int b = blockIdx.x, run = blockIdx.y;
    if (active && !active[b]) return;
    const int lane = threadIdx.x, N = s.N;
    const int i0 = P.seg_bounds[2 * run], i1 = P.seg_bounds[2 * run + 1], NS = i1 - i0;
    double* tmp = xs + seg_len_max + 96;
    double* z = tmp + 32;
    double* ring = z + 32;
    const int* meta = s.start;
    const int *m_start = meta, *m_diag = meta + N, *m_cls = meta + 7 * N, *m_pkF = meta + 8 * N, *m_szF = meta + 9 * N;
    const double* pk = pk_all + (size_t)b * pk_stride;
    const int base = m_start[i0], len = m_start[i1 - 1] + m_diag[i1 - 1] - base;
    double* x = X + (size_t)b * s.n + base;
    const int Dsep = run > 0 ? m_cls[i0 - 1] : 0;
    for (int q = 0; q < MSP_PF; q++) { if (q < NS) msw_issue(pk + m_pkF[i0 + q], m_szF[i0 + q], ring + (size_t)(q % MSP_PF) * slot_doubles, lane); msw_cp_commit(); }
    for (int e = lane; e < seg_len_max + 96; e += 32) xs[e] = e < len ? x[e] : 0.0;
    z[lane] = 0.0;
    for (int t = 0; t < NS; t++) {
        const int i = i0 + t;
        msw_cp_wait<MSP_PF - 1>();
        __syncwarp();
        const int d = m_diag[i], st = m_start[i] - base, D = m_cls[i];
        const int PD = t > 0 ? m_cls[i - 1] : 0, pst = t > 0 ? m_start[i - 1] - base : 0;
        const double* pkt = ring + (size_t)(t % MSP_PF) * slot_doubles;
        if (D == 16) msw_fwd_stage<16>(pkt, PD, xs, st, pst, d, 0, tmp, nullptr, lane);
        else if (D == 8) msw_fwd_stage<8>(pkt, PD, xs, st, pst, d, 0, tmp, nullptr, lane);
        else msw_fwd_stage<32>(pkt, PD, xs, st, pst, d, 0, tmp, nullptr, lane);
        if (run > 0) {                               // z += Y_i^T y_i (off the dependent chain)
            const double* YT = pkt + msp_yF(m_cls, i);
            if (lane < 32) tmp[lane] = lane < d ? xs[st + lane] : 0.0;       // zero-padded copy: the padded columns of Y^T are zero, the entries of xs behind the stage are not
            __syncwarp();
            double acc;
            if (Dsep == 16) { acc = msw_matvec_dyn<16>(YT, tmp, D, lane % 16, lane / 16); acc = msw_reduce_h<16>(acc); if (lane < 16) z[lane] += acc; }
            else if (Dsep == 8) { acc = msw_matvec_dyn<8>(YT, tmp, D, lane % 8, lane / 8); acc = msw_reduce_h<8>(acc); if (lane < 8) z[lane] += acc; }
            else { acc = msw_matvec_dyn<32>(YT, tmp, D, lane, 0); z[lane] += acc; }
        }
        __syncwarp();
        const int jn = t + MSP_PF;
        if (jn < NS) msw_issue(pk + m_pkF[i0 + jn], m_szF[i0 + jn], ring + (size_t)(jn % MSP_PF) * slot_doubles, lane);
        msw_cp_commit();
    }
    msw_cp_wait<0>();
    __syncwarp();
    for (int e = lane; e < len; e += 32) x[e] = xs[e];
    zbuf[((size_t)b * P.K + run) * 32 + lane] = z[lane];
}

// reduced right-hand side of separator k: x(g) - (B L^-T)(last stage of the run on the left) y(last) - Y^T y (run on the right); warp per separator
__global__ void msp_gather_kernel(MsDev s, MsPart P, const double* __restrict__ pk_all, size_t pk_stride, const double* __restrict__ X, const double* __restrict__ zbuf,
                                  double* __restrict__ xr, const int* __restrict__ active) {
    const int b = blockIdx.x;
    if (active && !active[b]) return;
    const int lane = threadIdx.x & 31, k = blockIdx.y * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (k >= P.K - 1) return;
    const int g = P.sep[k], il = g - 1;
    const int d = s.diag[g], D = s.cls[g], dl = s.diag[il], PD = s.cls[il];
    const double* pk = pk_all + (size_t)b * pk_stride;
    const double* x = X + (size_t)b * s.n;
    const double* Bs = pk + s.pkF[g] + D * D;          // (B L^-T)(il) as [D x PD], column-major
    if (lane < d) {
        double acc = 0.0;
        for (int q = 0; q < dl; q++) acc += Bs[lane + q * D] * x[s.start[il] + q];
        xr[(size_t)b * P.rn + P.rstart[k] + lane] = x[s.start[g] + lane] - acc - zbuf[((size_t)b * P.K + (k + 1)) * 32 + lane];
    }
    (void)PD;
}

__global__ void __launch_bounds__(32) msp_bwd_kernel(MsDev s, MsPart P, int slot_doubles, int seg_len_max, const double* __restrict__ pk_all, size_t pk_stride,
                                                     double* __restrict__ X, const double* __restrict__ xr_all, const int* __restrict__ active) {
    extern __shared__ __align__(16) double xs[];
    const int b = blockIdx.x, run = blockIdx.y;
    if (active && !active[b]) return;
    const int lane = threadIdx.x, N = s.N;
    const int i0 = P.seg_bounds[2 * run], i1 = P.seg_bounds[2 * run + 1], NS = i1 - i0;
    double* tmp = xs + seg_len_max + 96;
    double* xl = tmp + 32;
    double* ring = xl + 32;
    const int* meta = s.start;
    const int *m_start = meta, *m_diag = meta + N, *m_cls = meta + 7 * N, *m_pkB = meta + 10 * N, *m_szB = meta + 11 * N;
    const double* pk = pk_all + (size_t)b * pk_stride;
    const double* xr = xr_all + (size_t)b * P.rn;
    const int base = m_start[i0], len = m_start[i1 - 1] + m_diag[i1 - 1] - base;
    double* x = X + (size_t)b * s.n + base;
    const int Dsep = run > 0 ? m_cls[i0 - 1] : 0;
    for (int q = 0; q < MSP_PF; q++) { const int t = NS - 1 - q; if (t >= 0) msw_issue(pk + m_pkB[i0 + t], m_szB[i0 + t], ring + (size_t)(t % MSP_PF) * slot_doubles, lane); msw_cp_commit(); }
    for (int e = lane; e < seg_len_max + 96; e += 32) xs[e] = e < len ? x[e] : 0.0;
    __syncwarp();
    if (run + 1 < P.K) {                              // solution of the separator on the right sits where the next stage's x is read; it also goes back to X
        const int g = P.sep[run], d = m_diag[g];
        if (lane < d) { const double v = xr[P.rstart[run] + lane]; xs[len + lane] = v; X[(size_t)b * s.n + m_start[g] + lane] = v; }
    }
    { const int dl = run > 0 ? m_diag[i0 - 1] : 0; xl[lane] = lane < dl ? xr[P.rstart[run - (run > 0)] + lane] : 0.0; }
    __syncwarp();
    for (int t = NS - 1; t >= 0; t--) {
        const int i = i0 + t;
        msw_cp_wait<MSP_PF - 1>();
        __syncwarp();
        const int d = m_diag[i], st = m_start[i] - base, D = m_cls[i];
        const int ND = (i + 2 < N) ? m_cls[i + 1] : 0, nst = st + d;
        const double* pkt = ring + (size_t)(t % MSP_PF) * slot_doubles;
        if (run > 0) {                               // y_i -= Y_i x(separator on the left)
            const double* Y = pkt + msp_yB(m_cls, i, N);
            double acc;
            if (D == 16) { acc = msw_matvec_dyn<16>(Y, xl, Dsep, lane % 16, lane / 16); acc = msw_reduce_h<16>(acc); if (lane < 16 && lane < d) xs[st + lane] -= acc; }
            else if (D == 8) { acc = msw_matvec_dyn<8>(Y, xl, Dsep, lane % 8, lane / 8); acc = msw_reduce_h<8>(acc); if (lane < 8 && lane < d) xs[st + lane] -= acc; }
            else { acc = msw_matvec_dyn<32>(Y, xl, Dsep, lane, 0); if (lane < d) xs[st + lane] -= acc; }
            __syncwarp();
        }
        if (D == 16) msw_bwd_stage<16>(pkt, ND, xs, st, nst, d, 0, s.n, tmp, lane);
        else if (D == 8) msw_bwd_stage<8>(pkt, ND, xs, st, nst, d, 0, s.n, tmp, lane);
        else msw_bwd_stage<32>(pkt, ND, xs, st, nst, d, 0, s.n, tmp, lane);
        const int jn = t - MSP_PF;
        if (jn >= 0) msw_issue(pk + m_pkB[i0 + jn], m_szB[i0 + jn], ring + (size_t)(jn % MSP_PF) * slot_doubles, lane);
        msw_cp_commit();
    }
    msw_cp_wait<0>();
    __syncwarp();
    for (int e = lane; e < len; e += 32) x[e] = xs[e];
}

}  // namespace b200
