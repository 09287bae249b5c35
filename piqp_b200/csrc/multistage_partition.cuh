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

// run-local copy of the meta block: arrays of stride ml holding stages ib .. ib + ml - 1 (the separator on the left, the run, the separator
// on the right).  A meta read from global memory would put an L2 round trip on every stage of the dependent chains.
struct MspMeta {
    const int* m; int ml, ib;
    __device__ __forceinline__ int at(int a, int i) const { return m[a * ml + (i - ib)]; }
    __device__ __forceinline__ int start(int i) const { return at(0, i); }
    __device__ __forceinline__ int diag(int i) const { return at(1, i); }
    __device__ __forceinline__ int off(int i) const { return at(2, i); }
    __device__ __forceinline__ int offD(int i) const { return at(3, i); }
    __device__ __forceinline__ int offB(int i) const { return at(4, i); }
    __device__ __forceinline__ int cls(int i) const { return (i < ib) ? 0 : at(7, i); }
    __device__ __forceinline__ int pkF(int i) const { return at(8, i); }
    __device__ __forceinline__ int szF(int i) const { return at(9, i); }
    __device__ __forceinline__ int pkB(int i) const { return at(10, i); }
    __device__ __forceinline__ int szB(int i) const { return at(11, i); }
    __device__ __forceinline__ int yF(int i) const { const int D = cls(i); return D * D + D * (i > 0 ? cls(i - 1) : 0); }
    __device__ __forceinline__ int yB(int i, int N) const { const int D = cls(i); return D * D + D * ((i + 2 < N) ? cls(i + 1) : 0); }
};
constexpr int MSP_META_MAX = 40;      // stages per run + 2 that the run-local meta copy can hold (host checks)
__device__ __forceinline__ MspMeta msp_load_meta(int* dst, const MsDev& s, int i0, int i1, int lane, int nth) {
    MspMeta M;
    M.ib = i0 > 0 ? i0 - 1 : 0;
    const int ie = min(i1, s.N - 1);                 // inclusive: the stage after the run (a separator, or the arrow entry)
    M.ml = ie - M.ib + 1; M.m = dst;
    for (int e = lane; e < MS_META * M.ml; e += nth) { const int a = e / M.ml, i = e - a * M.ml; dst[e] = s.start[a * s.N + M.ib + i]; }
    return M;
}

// acc(8 ri + gq, 8 ci + 2 tq + e) += sum_k A(row, k) B(k, col): A at As[row + k * lda] (column-major), B at Bs[k * ldb + col] (row-major), K multiple of 4
__device__ __forceinline__ void msp_dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
// SNR / SNC > 0: static tile counts (the loops collapse to SNR x SNC DMMAs per k-step, no predicates); 0: the dynamic nr / nc.  The common stage
// (class 16 against a class-16 separator: 2 x 2 tiles) used to run through the dynamic form: 1 060 instructions per stage, 2.9 M ISETP per launch.
template <int SNR, int SNC>
__device__ __forceinline__ void msp_mma_t(double (&acc)[4][4][2], const double* As, int lda, const double* Bs, int ldb, int K, int nr, int nc, int lane) {
    const int gq = lane >> 2, tq = lane & 3;
    constexpr int RI = SNR > 0 ? SNR : 4, CI = SNC > 0 ? SNC : 4;
    for (int k0 = 0; k0 < K; k0 += 4) {
        double a[RI], b[CI];
#pragma unroll
        for (int ri = 0; ri < RI; ri++) a[ri] = (SNR > 0 || ri < nr) ? As[ri * 8 + gq + (k0 + tq) * lda] : 0.0;
#pragma unroll
        for (int ci = 0; ci < CI; ci++) b[ci] = (SNC > 0 || ci < nc) ? Bs[(k0 + tq) * ldb + ci * 8 + gq] : 0.0;
#pragma unroll
        for (int ri = 0; ri < RI; ri++)
#pragma unroll
            for (int ci = 0; ci < CI; ci++) if ((SNR > 0 || ri < nr) && (SNC > 0 || ci < nc)) msp_dmma(acc[ri][ci][0], acc[ri][ci][1], a[ri], b[ci]);
    }
}
__device__ __forceinline__ void msp_mma(double (&acc)[4][4][2], const double* As, int lda, const double* Bs, int ldb, int K, int nr, int nc, int lane) {
    msp_mma_t<0, 0>(acc, As, lda, Bs, ldb, K, nr, nc, lane);
}
template <int SNR, int SNC>
__device__ __forceinline__ void msp_zero_t(double (&acc)[4][4][2]) {
    constexpr int RI = SNR > 0 ? SNR : 4, CI = SNC > 0 ? SNC : 4;
#pragma unroll
    for (int ri = 0; ri < RI; ri++)
#pragma unroll
        for (int ci = 0; ci < CI; ci++) { acc[ri][ci][0] = 0.0; acc[ri][ci][1] = 0.0; }
}
__device__ __forceinline__ void msp_zero(double (&acc)[4][4][2]) { msp_zero_t<0, 0>(acc); }
constexpr int MSP_LDY = 36;           // row stride (doubles) of the 32 x 32 spike tiles in shared memory

// one stage of the spike recurrence (see msp_spike_kernel): V = -(B L^-T)(i-1) Y(i-1), Y(i) = inv(L_i) V -> Yc / the solve packets, G += Y(i)^T Y(i)
template <int SNR, int SNC>
__device__ __forceinline__ void msp_spike_stage(int t, const double* pkt, int D, int PD, int nr, int nc, int Dsep, const double* Yp, double* Yc, double* Vs,
                                                double* Yb, double* YTb, double (&G)[4][4][2], int lane) {
    const int gq = lane >> 2, tq = lane & 3;
    constexpr int RI = SNR > 0 ? SNR : 4, CI = SNC > 0 ? SNC : 4;
    double acc[4][4][2];
    if (t > 0) {                                 // V = -(B L^-T)(i-1) Y(i-1)
        msp_zero_t<SNR, SNC>(acc);
        msp_mma_t<SNR, SNC>(acc, pkt + D * D, D, Yp, MSP_LDY, PD, nr, nc, lane);
#pragma unroll
        for (int ri = 0; ri < RI; ri++)
#pragma unroll
            for (int ci = 0; ci < CI; ci++) if ((SNR > 0 || ri < nr) && (SNC > 0 || ci < nc)) {
                double* v = Vs + (ri * 8 + gq) * MSP_LDY + ci * 8 + 2 * tq;
                v[0] = -acc[ri][ci][0]; v[1] = -acc[ri][ci][1];
            }
    }
    __syncwarp();
    msp_zero_t<SNR, SNC>(acc);                   // Y(i) = inv(L_i) V   (rows >= d of the packet's inverse are zero)
    msp_mma_t<SNR, SNC>(acc, pkt, D, Vs, MSP_LDY, D, nr, nc, lane);
#pragma unroll
    for (int ri = 0; ri < RI; ri++)
#pragma unroll
        for (int ci = 0; ci < CI; ci++) if ((SNR > 0 || ri < nr) && (SNC > 0 || ci < nc)) {
            const int r = ri * 8 + gq, c = ci * 8 + 2 * tq;
            double* y = Yc + r * MSP_LDY + c;
            y[0] = acc[ri][ci][0]; y[1] = acc[ri][ci][1];
            Yb[r + c * D] = acc[ri][ci][0]; Yb[r + (c + 1) * D] = acc[ri][ci][1];      // Y   [D x Dsep], column-major
            *reinterpret_cast<double2*>(YTb + c + r * Dsep) = make_double2(acc[ri][ci][0], acc[ri][ci][1]);      // Y^T [Dsep x D]
        }
    __syncwarp();
    msp_mma_t<SNC, SNC>(G, Yc, MSP_LDY, Yc, MSP_LDY, D, nc, nc, lane);          // G += Y(i)^T Y(i): A(j1, r) = Yc[r][j1]
}

// ---- 2 + 3. spikes and the reduced system: ONE WARP per (QP, run s >= 1).  The spike Y = L^-1 K[run, g] has d(g) <= 32 columns; a stage
// is two small matrix products on the FP64 tensor pipe (DMMA m8n8k4):  V = rhs - (B L^-T)(i-1) Y(i-1) ,  Y(i) = inv(L_i) V ,  plus
// the Gram update G += Y(i)^T Y(i).  Forward packets stream through a per-warp cp.async ring; Y / Y^T go into the solve packets.  At the
// end the warp emits the reduced blocks of separator k = run - 1:
//     D~_k = D(g_k) + carry(run on the left) - Y^T Y ,   B~_k = -(B L^-T)(last stage of this run) Y(last stage)
// smem: Yp, Yc, Vs [32][MSP_LDY] | meta | ring[MSP_PF][slot_spike]
__global__ void __launch_bounds__(32) msp_spike_kernel(MsDev s, MsPart P, int slot_doubles, const double* __restrict__ fac_all,
                                                       double* __restrict__ pk_all, size_t pk_stride, const double* __restrict__ carry_all,
                                                       double* __restrict__ rfac_all, const int* __restrict__ active) {
    extern __shared__ __align__(16) double sp_sm[];
    const int b = blockIdx.x;
    if (active && !active[b]) return;
    const int run = blockIdx.y + 1, k = run - 1;
    const int lane = threadIdx.x, gq = lane >> 2, tq = lane & 3;
    const int N = s.N;
    const int i0 = P.seg_bounds[2 * run], i1 = P.seg_bounds[2 * run + 1], g = i0 - 1, NS = i1 - i0;
    double* Yp = sp_sm;
    double* Yc = Yp + 32 * MSP_LDY;
    double* Vs = Yc + 32 * MSP_LDY;
    int* metab = reinterpret_cast<int*>(Vs + 32 * MSP_LDY);
    double* ring = Vs + 32 * MSP_LDY + (MS_META * MSP_META_MAX + 1) / 2;
    const MspMeta M = msp_load_meta(metab, s, i0, i1, lane, 32);
    __syncwarp();
    const double* fac = fac_all + (size_t)b * s.total;
    double* pk = pk_all + (size_t)b * pk_stride;
    const int dg = M.diag(g), og = M.off(g), Dsep = M.cls(g), nc = Dsep / 8;
    // A packet larger than a ring slot (a stage of a bigger class than the common one, e.g. the 28-variable last stage of the MPC
    // problems next to 16-variable stages) is read from global memory instead: sizing the slots for it would halve the number of
    // resident CTAs for every run.
    auto psize = [&](int i) { const int D = M.cls(i); return D * D + D * (i > 0 ? M.cls(i - 1) : 0); };
    auto issue = [&](int t) {                        // inv(L_i) | (B L^-T)(i-1) of the forward packet
        const int i = i0 + t, sz = psize(i);
        if (sz <= slot_doubles) msw_issue(pk + M.pkF(i), sz, ring + (size_t)(t % MSP_PF) * slot_doubles, lane);
    };
    for (int q = 0; q < MSP_PF; q++) { if (q < NS) issue(q); msw_cp_commit(); }
    // first stage: V = K[first stage, g] = B(g): rows = the first og variables of the run, columns = the dg variables of the separator
    for (int e = lane; e < 32 * MSP_LDY; e += 32) { Yp[e] = 0.0; Yc[e] = 0.0; Vs[e] = 0.0; }
    __syncwarp();
    for (int e = lane; e < og * dg; e += 32) { const int r = e % og, j = e / og; Vs[r * MSP_LDY + j] = __ldg(fac + M.offB(g) + e); }
    double G[4][4][2];
    msp_zero(G);
    for (int t = 0; t < NS; t++) {
        const int i = i0 + t, d = M.diag(i), D = M.cls(i), nr = D / 8;
        const int PD = t > 0 ? M.cls(i - 1) : 0;
        msw_cp_wait<MSP_PF - 1>();
        __syncwarp();
        const double* pkt = psize(i) <= slot_doubles ? ring + (size_t)(t % MSP_PF) * slot_doubles : pk + M.pkF(i);
        if (nr == 2 && nc == 2 && (t == 0 || PD == 16)) msp_spike_stage<2, 2>(t, pkt, D, PD, nr, nc, Dsep, Yp, Yc, Vs, pk + M.pkB(i) + M.yB(i, N), pk + M.pkF(i) + M.yF(i), G, lane);
        else msp_spike_stage<0, 0>(t, pkt, D, PD, nr, nc, Dsep, Yp, Yc, Vs, pk + M.pkB(i) + M.yB(i, N), pk + M.pkF(i) + M.yF(i), G, lane);
        { const int tn = t + MSP_PF; if (tn < NS) issue(tn); msw_cp_commit(); }
        double* tmpp = Yp; Yp = Yc; Yc = tmpp;       // Y(i) becomes the previous stage's spike
        (void)d;
    }
    msw_cp_wait<0>();
    __syncwarp();
    // ---- reduced blocks of separator k (lower triangle of D~; the upper part stays zero like in the assembled blocks of the chain)
    double* rfac = rfac_all + (size_t)b * P.rtotal;
    {
        const int ol_ok = s.off[g - 1];             // coupling rows of the last stage of the run on the left (outside the run-local meta copy)
        const double* carry = carry_all + ((size_t)b * P.K + k) * 1024;      // Schur complement the run on the LEFT left on its coupling rows
#pragma unroll
        for (int ri = 0; ri < 4; ri++)
#pragma unroll
            for (int ci = 0; ci < 4; ci++) if (ri < nc && ci < nc)
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const int r = ri * 8 + gq, c = ci * 8 + 2 * tq + e;
                    if (r < dg && c < dg) {
                        double v = 0.0;
                        if (r >= c) {
                            v = fac[M.offD(g) + r + (size_t)c * dg];
                            if (r < ol_ok) v += carry[r + 32 * c];
                            v -= G[ri][ci][e];
                        }
                        rfac[P.roffD[k] + r + c * dg] = v;
                    }
                }
    }
    if (k + 1 < P.K - 1) {                           // B~_k = -(B L^-T)(last stage) Y(last): the separator on the right holds (B L^-T)(il) in its forward packet
        const int il = i1 - 1, gn = i1, o2 = M.off(il), Dl = M.cls(il), Dn = M.cls(gn);
        const double* Bh = pk + M.pkF(gn) + Dn * Dn;                         // [Dn x Dl] column-major, rows >= o2 zero (written by the chain kernel of this run)
        double acc[4][4][2];
        msp_zero(acc);
        msp_mma(acc, Bh, Dn, Yp, MSP_LDY, Dl, Dn / 8, nc, lane);             // Yp = Y(last) after the final swap
#pragma unroll
        for (int ri = 0; ri < 4; ri++)
#pragma unroll
            for (int ci = 0; ci < 4; ci++) if (ri < Dn / 8 && ci < nc)
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const int r = ri * 8 + gq, c = ci * 8 + 2 * tq + e;
                    if (r < o2 && c < dg) rfac[P.roffB[k] + r + c * o2] = -acc[ri][ci][e];
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
    int* metab = reinterpret_cast<int*>(z + 32);
    double* ring = z + 32 + (MS_META * MSP_META_MAX + 1) / 2;
    const MspMeta M = msp_load_meta(metab, s, i0, i1, lane, 32);
    __syncwarp();
    const double* pk = pk_all + (size_t)b * pk_stride;
    const int base = M.start(i0), len = M.start(i1 - 1) + M.diag(i1 - 1) - base;
    double* x = X + (size_t)b * s.n + base;
    const int Dsep = run > 0 ? M.cls(i0 - 1) : 0;
    for (int q = 0; q < MSP_PF; q++) { if (q < NS && M.szF(i0 + q) <= slot_doubles) msw_issue(pk + M.pkF(i0 + q), M.szF(i0 + q), ring + (size_t)(q % MSP_PF) * slot_doubles, lane); msw_cp_commit(); }
    for (int e = lane; e < seg_len_max + 96; e += 32) xs[e] = e < len ? x[e] : 0.0;
    z[lane] = 0.0;
    for (int t = 0; t < NS; t++) {
        const int i = i0 + t;
        msw_cp_wait<MSP_PF - 1>();
        __syncwarp();
        const int d = M.diag(i), st = M.start(i) - base, D = M.cls(i);
        const int PD = t > 0 ? M.cls(i - 1) : 0, pst = t > 0 ? M.start(i - 1) - base : 0;
        const double* pkt = M.szF(i) <= slot_doubles ? ring + (size_t)(t % MSP_PF) * slot_doubles : pk + M.pkF(i);      // oversized packet: straight from global memory
        if (D == 16) msw_fwd_stage<16>(pkt, PD, xs, st, pst, d, 0, tmp, nullptr, lane);
        else if (D == 8) msw_fwd_stage<8>(pkt, PD, xs, st, pst, d, 0, tmp, nullptr, lane);
        else msw_fwd_stage<32>(pkt, PD, xs, st, pst, d, 0, tmp, nullptr, lane);
        if (run > 0) {                               // z += Y_i^T y_i (off the dependent chain)
            const double* YT = pkt + M.yF(i);
            if (lane < 32) tmp[lane] = lane < d ? xs[st + lane] : 0.0;       // zero-padded copy: the padded columns of Y^T are zero, the entries of xs behind the stage are not
            __syncwarp();
            double acc;
            if (Dsep == 16) { acc = msw_matvec_dyn<16>(YT, tmp, D, lane % 16, lane / 16); acc = msw_reduce_h<16>(acc); if (lane < 16) z[lane] += acc; }
            else if (Dsep == 8) { acc = msw_matvec_dyn<8>(YT, tmp, D, lane % 8, lane / 8); acc = msw_reduce_h<8>(acc); if (lane < 8) z[lane] += acc; }
            else { acc = msw_matvec_dyn<32>(YT, tmp, D, lane, 0); z[lane] += acc; }
        }
        __syncwarp();
        const int jn = t + MSP_PF;
        if (jn < NS && M.szF(i0 + jn) <= slot_doubles) msw_issue(pk + M.pkF(i0 + jn), M.szF(i0 + jn), ring + (size_t)(jn % MSP_PF) * slot_doubles, lane);
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
    int* metab = reinterpret_cast<int*>(xl + 32);
    double* ring = xl + 32 + (MS_META * MSP_META_MAX + 1) / 2;
    const MspMeta M = msp_load_meta(metab, s, i0, i1, lane, 32);
    __syncwarp();
    const double* pk = pk_all + (size_t)b * pk_stride;
    const double* xr = xr_all + (size_t)b * P.rn;
    const int base = M.start(i0), len = M.start(i1 - 1) + M.diag(i1 - 1) - base;
    double* x = X + (size_t)b * s.n + base;
    const int Dsep = run > 0 ? M.cls(i0 - 1) : 0;
    for (int q = 0; q < MSP_PF; q++) { const int t = NS - 1 - q; if (t >= 0 && M.szB(i0 + t) <= slot_doubles) msw_issue(pk + M.pkB(i0 + t), M.szB(i0 + t), ring + (size_t)(t % MSP_PF) * slot_doubles, lane); msw_cp_commit(); }
    for (int e = lane; e < seg_len_max + 96; e += 32) xs[e] = e < len ? x[e] : 0.0;
    __syncwarp();
    if (run + 1 < P.K) {                              // solution of the separator on the right sits where the next stage's x is read; it also goes back to X
        const int g = P.sep[run], d = M.diag(g);
        if (lane < d) { const double v = xr[P.rstart[run] + lane]; xs[len + lane] = v; X[(size_t)b * s.n + M.start(g) + lane] = v; }
    }
    { const int dl = run > 0 ? M.diag(i0 - 1) : 0; xl[lane] = lane < dl ? xr[P.rstart[run - (run > 0)] + lane] : 0.0; }
    __syncwarp();
    for (int t = NS - 1; t >= 0; t--) {
        const int i = i0 + t;
        msw_cp_wait<MSP_PF - 1>();
        __syncwarp();
        const int d = M.diag(i), st = M.start(i) - base, D = M.cls(i);
        const int ND = (i + 2 < N) ? M.cls(i + 1) : 0, nst = st + d;
        const double* pkt = M.szB(i) <= slot_doubles ? ring + (size_t)(t % MSP_PF) * slot_doubles : pk + M.pkB(i);
        if (run > 0) {                               // y_i -= Y_i x(separator on the left)
            const double* Y = pkt + M.yB(i, N);
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
        if (jn >= 0 && M.szB(i0 + jn) <= slot_doubles) msw_issue(pk + M.pkB(i0 + jn), M.szB(i0 + jn), ring + (size_t)(jn % MSP_PF) * slot_doubles, lane);
        msw_cp_commit();
    }
    msw_cp_wait<0>();
    __syncwarp();
    for (int e = lane; e < len; e += 32) x[e] = xs[e];
}


// =====================================================================================================
// Fused solve: ONE launch per backend solve instead of msp_fwd / msp_gather / msw_solve (reduced) / msp_bwd.
//
// The four kernels above are single-warp instruction streams of ~300 instructions per stage at ~10 cycles per instruction (ncu source view,
// profiles/): their time is the length of that stream plus, per launch, a start-up of meta block -> first packets -> x (three dependent
// round trips to L2) and the launch gap.  Here a CTA owns one QP: warp r < K walks run r and keeps its slice of x in shared memory across
// the forward and the backward sweep; warp K + r is the run's HELPER and takes everything that is not on the dependent chain off it:
//   forward : z += Y_i^T y_i  as soon as the chain warp has published y_i (a counter in shared memory), Y^T read from global memory,
//   backward: y_i -= Y_i x(separator on the left), running AHEAD of the chain warp, which waits for the counter before it touches stage i.
// The chain warp's packets then shrink to inv(L_i) | (B L^-T) (the spikes stay out of its cp.async ring), and its common stage (class 16
// next to class 16) runs with static shapes.  Between the sweeps warp k assembles the reduced right-hand side of separator k and warp 0
// solves the reduced chain (msw_solve_body).  __syncthreads() separates the three phases; x / zbuf / xred cross them through global memory
// exactly as between the separate kernels.
// smem per CTA: K x [ xs[seg_len_max + 96] | tmp[32] | htmp[32] | xl[32] | ring[MSF_PF][slot] | meta | 2 counters ]  +  workspace of msw_solve_body
// =====================================================================================================
constexpr int MSF_PF = 4;
__device__ __forceinline__ void msf_publish(volatile int* flag, int v, int lane) { __syncwarp(); if (lane == 0) { __threadfence_block(); *flag = v; } }
__device__ __forceinline__ void msf_wait(volatile int* flag, int v) { while (*flag < v) { } __threadfence_block(); __syncwarp(); }
__device__ __forceinline__ void msf_pair_sync(int run) { asm volatile("bar.sync %0, 64;" ::"r"(1 + run) : "memory"); }
__device__ __forceinline__ void msf_issue(const double* src, int sz, double* slot, int lane) {
    if (sz == 512) {
#pragma unroll
        for (int j = 0; j < 8; j++) msw_cp_async16(slot + 2 * lane + 64 * j, src + 2 * lane + 64 * j);
    } else msw_issue(src, sz, slot, lane);
}
// the common stage with static shapes: class 16 behind / in front of class 16, packet in shared memory
__device__ __forceinline__ void msf_fwd_stage16(const double* pkt, double* xs, int st, int pst, int d, double* tmp, int lane) {
    const int r = lane & 15, h = lane >> 4;
    double acc = msw_matvec<16, 16>(pkt + 256, xs + pst, r, h);
    acc = msw_reduce_h<16>(acc);
    if (h == 0) tmp[r] = xs[st + r] - acc;
    __syncwarp();
    acc = msw_matvec<16, 16>(pkt, tmp, r, h);
    acc = msw_reduce_h<16>(acc);
    if (h == 0 && r < d) xs[st + r] = acc;
    __syncwarp();
}
__device__ __forceinline__ void msf_bwd_stage16(const double* pkt, double* xs, int st, int nst, int d, double* tmp, int lane) {
    const int r = lane & 15, h = lane >> 4;
    double acc = msw_matvec<16, 16>(pkt + 256, xs + nst, r, h);
    acc = msw_reduce_h<16>(acc);
    if (h == 0) tmp[r] = (r < d) ? xs[st + r] - acc : 0.0;
    __syncwarp();
    acc = msw_matvec<16, 16>(pkt, tmp, r, h);
    acc = msw_reduce_h<16>(acc);
    if (h == 0 && r < d) xs[st + r] = acc;
    __syncwarp();
}

__global__ void __launch_bounds__(512) msp_solve_fused_kernel(MsDev s, MsPart P, MsDev rd, int slot_doubles, int seg_len_max, int run_doubles, int rslot_doubles,
                                                               const double* __restrict__ pk_all, size_t pk_stride, const double* __restrict__ rpk_all, size_t rpk_stride,
                                                               double* __restrict__ X, double* __restrict__ zbuf, double* __restrict__ xred_all, const int* __restrict__ active) {
    extern __shared__ __align__(16) double msf_sm[];
    const int b = blockIdx.x;
    if (active && !active[b]) return;
    const int K = P.K, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, N = s.N;
    const bool helper = warp >= K;
    const int run = helper ? warp - K : warp;
    const int i0 = P.seg_bounds[2 * run], i1 = P.seg_bounds[2 * run + 1], NS = i1 - i0;
    double* xs = msf_sm + (size_t)run * run_doubles;
    double* tmp = xs + seg_len_max + 96;
    double* htmp = tmp + 32;
    double* xl = htmp + 32;
    double* ring = xl + 32;
    int* metab = reinterpret_cast<int*>(ring + (size_t)MSF_PF * slot_doubles);
    volatile int* flags = metab + MS_META * MSP_META_MAX;
    MspMeta M;                                          // run-local meta block, loaded by the pair (warp r, warp K + r) together
    M.ib = i0 > 0 ? i0 - 1 : 0; M.ml = min(i1, N - 1) - M.ib + 1; M.m = metab;
    const int pl = (helper ? 32 : 0) + lane;
    for (int e = pl; e < MS_META * M.ml; e += 64) { const int a = e / M.ml, i = e - a * M.ml; metab[e] = s.start[a * N + M.ib + i]; }
    if (pl == 0) { flags[0] = 0; flags[1] = 0; }
    const double* pk = pk_all + (size_t)b * pk_stride;
    const double* xr = xred_all + (size_t)b * P.rn;
    msf_pair_sync(run);
    const int base = M.start(i0), len = M.start(i1 - 1) + M.diag(i1 - 1) - base;
    double* x = X + (size_t)b * s.n + base;
    const int Dsep = run > 0 ? M.cls(i0 - 1) : 0;
    auto szFc = [&](int i) { return M.yF(i); };             // chain part of the forward / backward packet (the spike blocks follow it)
    auto szBc = [&](int i) { return M.yB(i, N); };

    // ---------------- phase 1: forward sweep of every run ----------------
    if (!helper) {
        for (int q = 0; q < MSF_PF; q++) { if (q < NS && szFc(i0 + q) <= slot_doubles) msf_issue(pk + M.pkF(i0 + q), szFc(i0 + q), ring + (size_t)q * slot_doubles, lane); msw_cp_commit(); }
        for (int e = lane; e < seg_len_max + 96; e += 32) xs[e] = e < len ? x[e] : 0.0;
        for (int t = 0; t < NS; t++) {
            const int i = i0 + t;
            msw_cp_wait<MSF_PF - 1>();
            __syncwarp();
            const int d = M.diag(i), st = M.start(i) - base, D = M.cls(i);
            const int PD = t > 0 ? M.cls(i - 1) : 0, pst = t > 0 ? M.start(i - 1) - base : 0;
            const bool in_ring = szFc(i) <= slot_doubles;
            const double* rpkt = ring + (size_t)(t % MSF_PF) * slot_doubles;
            if (in_ring && D == 16 && PD == 16) msf_fwd_stage16(rpkt, xs, st, pst, d, tmp, lane);
            else {
                const double* pkt = in_ring ? rpkt : pk + M.pkF(i);
                if (D == 16) msw_fwd_stage<16>(pkt, PD, xs, st, pst, d, 0, tmp, nullptr, lane);
                else if (D == 8) msw_fwd_stage<8>(pkt, PD, xs, st, pst, d, 0, tmp, nullptr, lane);
                else msw_fwd_stage<32>(pkt, PD, xs, st, pst, d, 0, tmp, nullptr, lane);
            }
            if (run > 0) msf_publish(flags, t + 1, lane);
            const int jn = t + MSF_PF;
            if (jn < NS && szFc(i0 + jn) <= slot_doubles) msf_issue(pk + M.pkF(i0 + jn), szFc(i0 + jn), ring + (size_t)(jn % MSF_PF) * slot_doubles, lane);
            msw_cp_commit();
        }
        msw_cp_wait<0>();
        __syncwarp();
        for (int e = lane; e < len; e += 32) x[e] = xs[e];          // y of the run: the gather below reads the last stage's part
    } else if (run > 0) {
        double z = 0.0;
        for (int t = 0; t < NS; t++) {
            const int i = i0 + t;
            const int d = M.diag(i), st = M.start(i) - base, D = M.cls(i);
            const double* YT = pk + M.pkF(i) + M.yF(i);
            msf_wait(flags, t + 1);
            htmp[lane] = lane < d ? xs[st + lane] : 0.0;          // zero-padded copy: the padded columns of Y^T are zero, the entries of xs behind the stage are not
            __syncwarp();
            double acc;
            if (Dsep == 16) { acc = msw_matvec_dyn<16>(YT, htmp, D, lane % 16, lane / 16); acc = msw_reduce_h<16>(acc); }
            else if (Dsep == 8) { acc = msw_matvec_dyn<8>(YT, htmp, D, lane % 8, lane / 8); acc = msw_reduce_h<8>(acc); }
            else acc = msw_matvec_dyn<32>(YT, htmp, D, lane, 0);
            z += acc;
            __syncwarp();
        }
        zbuf[((size_t)b * K + run) * 32 + lane] = lane < Dsep ? z : 0.0;
    }
    __syncthreads();

    // ---------------- phase 2: reduced right-hand side (warp k: separator k), reduced chain (warp 0) ----------------
    if (warp < K - 1) {
        const int k = warp, g = P.sep[k], il = g - 1;
        const int d = s.diag[g], D = s.cls[g], dl = s.diag[il];
        const double* xg = X + (size_t)b * s.n;
        const double* Bs = pk + s.pkF[g] + D * D;          // (B L^-T)(il) as [D x PD], column-major
        if (lane < d) {
            double acc = 0.0;
            for (int q = 0; q < dl; q++) acc += Bs[lane + q * D] * xg[s.start[il] + q];
            xred_all[(size_t)b * P.rn + P.rstart[k] + lane] = xg[s.start[g] + lane] - acc - zbuf[((size_t)b * K + (k + 1)) * 32 + lane];
        }
    }
    __syncthreads();
    if (warp == 0) msw_solve_body(rd, rslot_doubles, nullptr, rpk_all + (size_t)b * rpk_stride, xred_all + (size_t)b * P.rn, msf_sm + (size_t)K * run_doubles, lane);
    __syncthreads();

    // ---------------- phase 3: backward sweep of every run (xs still holds the run's y) ----------------
    if (!helper) {
        for (int q = 0; q < MSF_PF; q++) { const int t = NS - 1 - q; if (t >= 0 && szBc(i0 + t) <= slot_doubles) msf_issue(pk + M.pkB(i0 + t), szBc(i0 + t), ring + (size_t)(t % MSF_PF) * slot_doubles, lane); msw_cp_commit(); }
        if (run + 1 < K) {                              // solution of the separator on the right sits where the next stage's x is read; it also goes back to X
            const int g = P.sep[run], d = M.diag(g);
            if (lane < d) { const double v = xr[P.rstart[run] + lane]; xs[len + lane] = v; X[(size_t)b * s.n + M.start(g) + lane] = v; }
        }
        __syncwarp();
        for (int t = NS - 1; t >= 0; t--) {
            const int i = i0 + t;
            msw_cp_wait<MSF_PF - 1>();
            __syncwarp();
            const int d = M.diag(i), st = M.start(i) - base, D = M.cls(i);
            const int ND = (i + 2 < N) ? M.cls(i + 1) : 0, nst = st + d;
            if (run > 0) msf_wait(flags + 1, NS - t);          // the helper has subtracted Y_i x(separator on the left) from this stage
            const bool in_ring = szBc(i) <= slot_doubles;
            const double* rpkt = ring + (size_t)(t % MSF_PF) * slot_doubles;
            if (in_ring && D == 16 && ND == 16) msf_bwd_stage16(rpkt, xs, st, nst, d, tmp, lane);
            else {
                const double* pkt = in_ring ? rpkt : pk + M.pkB(i);
                if (D == 16) msw_bwd_stage<16>(pkt, ND, xs, st, nst, d, 0, s.n, tmp, lane);
                else if (D == 8) msw_bwd_stage<8>(pkt, ND, xs, st, nst, d, 0, s.n, tmp, lane);
                else msw_bwd_stage<32>(pkt, ND, xs, st, nst, d, 0, s.n, tmp, lane);
            }
            const int jn = t - MSF_PF;
            if (jn >= 0 && szBc(i0 + jn) <= slot_doubles) msf_issue(pk + M.pkB(i0 + jn), szBc(i0 + jn), ring + (size_t)(jn % MSF_PF) * slot_doubles, lane);
            msw_cp_commit();
        }
        msw_cp_wait<0>();
        __syncwarp();
        for (int e = lane; e < len; e += 32) x[e] = xs[e];
    } else if (run > 0) {
        const int dl = M.diag(i0 - 1);
        xl[lane] = lane < dl ? xr[P.rstart[run - 1] + lane] : 0.0;
        __syncwarp();
        for (int t = NS - 1; t >= 0; t--) {
            const int i = i0 + t;
            const int d = M.diag(i), st = M.start(i) - base, D = M.cls(i);
            const double* Y = pk + M.pkB(i) + M.yB(i, N);
            double acc;
            if (D == 16) { acc = msw_matvec_dyn<16>(Y, xl, Dsep, lane % 16, lane / 16); acc = msw_reduce_h<16>(acc); if (lane < 16 && lane < d) xs[st + lane] -= acc; }
            else if (D == 8) { acc = msw_matvec_dyn<8>(Y, xl, Dsep, lane % 8, lane / 8); acc = msw_reduce_h<8>(acc); if (lane < 8 && lane < d) xs[st + lane] -= acc; }
            else { acc = msw_matvec_dyn<32>(Y, xl, Dsep, lane, 0); if (lane < d) xs[st + lane] -= acc; }
            msf_publish(flags + 1, NS - t, lane);
        }
    }
}

}  // namespace b200
