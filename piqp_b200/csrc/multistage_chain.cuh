// piqp_b200/csrc/multistage_chain.cuh -- latency-optimised chain kernels of the multistage backend.
//
// factor_kkt (multistage_kkt.hpp:1253-1352) and solve_llt_in_place (:1709-1816) are chains of N dependent small-block
// steps per QP, so their time is (stages) x (latency of one step).  When a stage's front (d pivot rows + o coupling rows
// + w arrow rows) has at most 32 rows -- BASELINE config 4: d=16, o=12, w=0 -- one WARP walks the chain:
//
//   msw_factor_kernel : lane r owns row r of the current front in REGISTERS (the stage is one partial dense Cholesky of a
//                       <=32x32 front, right-looking: potrf, the two trsm and the three syrk/gemm updates of :1289-1345
//                       in one sweep of d pivots); pivot columns are broadcast through shared memory (one LDS.128 per
//                       two multipliers, all loop bounds static), the Schur complement is carried to the next stage
//                       through a 32x33 smem tile.  A second warp trails one stage behind: it inverts L_i and writes
//                       the stage's SOLVE PACKETS.
//   solve packets     : per stage and direction one contiguous, 16-byte aligned, zero-padded record holding exactly what
//                       the substitution step needs, already in the orientation it is read in:
//                         forward  i : inv(L_i) [D x D] | B_{i-1} [D x PD] | E_i [w x D]
//                         backward i : inv(L_i)^T [D x D] | B_i^T [D x ND] | E_i^T [D x WP]
//                       with D = 8/16/32 >= d_i (class of the stage), PD / ND the class of the previous / next stage.
//   msw_solve_kernel  : one warp; packets are streamed into a shared-memory ring with 16-byte cp.async MSW_PF stages
//                       ahead of the dependent chain; each stage is two mat-vecs with fully static loops and 32/D lanes
//                       per row.
//
// Stages whose front exceeds 32 rows use the general shared-memory kernels in multistage_backend.cu.
#pragma once
#include "common.cuh"

namespace b200 {

struct MsDev {
    const int *start, *diag, *off, *offD, *offB, *offE, *offI;
    const int *cls, *pkF, *szF, *pkB, *szB;      // warp-chain only
    int N, w, n, total, total_inv, dmax, omax;
};
constexpr int MS_META = 12;                       // int arrays of length N in the meta block, in the order above

constexpr int MSW_LD = 33;    // odd leading dimension (doubles) of the carried Schur tile

__device__ __forceinline__ void msw_cp_async16(double* dst_smem, const double* src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void msw_cp_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void msw_cp_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// ------------------------------------------------------------------------------------------------------------------
// factor: 2 warps per QP.  warp 0 = chain, warp 1 = inverse of the pivot block + packet write-out (one stage behind).
// RP = static bound on ceil(front rows / 2).
// ------------------------------------------------------------------------------------------------------------------
struct MswFactorSmem {
    double L[2][32][32];      // pivot columns of the front after elimination: L[buf][k][lane]
    double rs[2][32];         // rsqrt(pivot)
    double S[32 * MSW_LD];    // carried Schur complement
};

template <int RP>
__global__ void __launch_bounds__(64) msw_factor_kernel(MsDev s, const double* __restrict__ fac_all, double* __restrict__ inv_all, double* __restrict__ pk_all,
                                                        size_t pk_stride, const int* __restrict__ active) {
    __shared__ __align__(16) MswFactorSmem sm;
    extern __shared__ int msw_meta[];                // MS_META * N ints
    const int b = blockIdx.x;
    if (active && !active[b]) return;
    const double* fac = fac_all + (size_t)b * s.total;
    double* Linv = inv_all + (size_t)b * s.total_inv;
    double* pk = pk_all + (size_t)b * pk_stride;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int N = s.N, w = s.w;
    const int nst = (w > 0) ? N : N - 1;          // the arrow corner D_N is one more stage (d = w, o = 0, no arrow rows)
    for (int e = threadIdx.x; e < MS_META * N; e += 64) msw_meta[e] = s.start[e];      // the meta block is contiguous, start[] first
    for (int e = threadIdx.x; e < 32 * MSW_LD; e += 64) sm.S[e] = 0.0;
    __syncthreads();
    const int* meta = msw_meta;
    const int *m_diag = meta + N, *m_off = meta + 2 * N, *m_offD = meta + 3 * N, *m_offB = meta + 4 * N, *m_offE = meta + 5 * N,
              *m_offI = meta + 6 * N, *m_cls = meta + 7 * N, *m_pkF = meta + 8 * N, *m_pkB = meta + 10 * N;

    if (warp == 0) {
        double nxt[2 * RP];
        // prefetch of the assembled blocks of a stage: lane r reads row r of [D_i; B_i; E_i] (column-major blocks)
        auto prefetch = [&](int i) {
            const int d = m_diag[i];
            const bool corner = (i == N - 1);
            const int o = (i + 2 < N) ? m_off[i] : 0, ww = corner ? 0 : w;
            const double* src = nullptr; int ld = 0;
            if (lane < d) { src = fac + m_offD[i] + lane; ld = d; }
            else if (lane < d + o) { src = fac + m_offB[i] + (lane - d); ld = o; }
            else if (lane < d + o + ww) { src = fac + m_offE[i] + (lane - d - o); ld = ww; }
#pragma unroll
            for (int c = 0; c < 2 * RP; c++) nxt[c] = (src && c < d) ? __ldg(src + (size_t)c * ld) : 0.0;
        };
        prefetch(0);
        int po = 0;                                 // coupling rows of the previous stage
        for (int i = 0; i < nst; i++) {
            const bool corner = (i == N - 1);
            const int d = m_diag[i], o = (i + 2 < N) ? m_off[i] : 0, ww = corner ? 0 : w;
            const int rows = d + o + ww;
            double v[2 * RP];
            // ---- assemble the front: prefetched K blocks + carried Schur complement
            {
                const bool isD = lane < d, isF = lane >= d + o && lane < rows;
                const int f = lane - d - o;
                const double* Srow = isD ? (sm.S + (corner ? po + lane : lane) * MSW_LD) : (isF ? sm.S + (po + f) * MSW_LD : sm.S);
#pragma unroll
                for (int c = 0; c < 2 * RP; c++) {
                    double x = nxt[c];
                    if (c < d) {
                        if (corner) { if (isD) x += Srow[po + c]; }
                        else if (c < po && ((isD && lane < po) || isF)) x += Srow[c];
                    } else if (isF && c >= d + o && c < rows) x = Srow[po + (c - d - o)];
                    v[c] = x;
                }
            }
            if (i + 1 < nst) prefetch(i + 1);
            double (*Lb)[32] = sm.L[i & 1];
            double* rsb = sm.rs[i & 1];
            // ---- d pivots, right-looking over the whole front (columns beyond `rows` compute on zeros / unused lanes)
#pragma unroll
            for (int k = 0; k < 2 * RP; k++) {
                if (k >= d) break;
                const double piv = __shfl_sync(0xffffffffu, v[k], k);
                const double r = rsqrt(piv);
                const double l = v[k] * r;            // lane k: piv * rsqrt(piv) = sqrt(piv)
                Lb[k][lane] = l;
                if (lane == 0) rsb[k] = r;
                __syncwarp();
                const double2* col = reinterpret_cast<const double2*>(&Lb[k][0]);
                double2 q[RP];
#pragma unroll
                for (int c2 = (k + 1) / 2; c2 < RP; c2++) q[c2] = col[c2];
#pragma unroll
                for (int c2 = (k + 1) / 2; c2 < RP; c2++) {
                    if (2 * c2 > k) v[2 * c2] -= l * q[c2].x;
                    v[2 * c2 + 1] -= l * q[c2].y;
                }
            }
            // ---- carry the Schur complement (rows/cols d..rows-1) to the next stage
            if (lane >= d && lane < rows) {
                double* Srow = sm.S + (lane - d) * MSW_LD;
#pragma unroll
                for (int c = 0; c < 2 * RP; c++) if (c >= d && c < rows) Srow[c - d] = v[c];
            }
            po = o;
            __syncthreads();                         // hand L columns of stage i to warp 1 (also orders S for the next stage)
        }
    } else {
        for (int i = 0; i < nst; i++) {
            __syncthreads();
            const bool corner = (i == N - 1);
            const int d = m_diag[i], o = (i + 2 < N) ? m_off[i] : 0, ww = corner ? 0 : w;
            const int rows = d + o + ww;
            const double (*Lb)[32] = sm.L[i & 1];
            const double* rsb = sm.rs[i & 1];
            // inv(L_i): lane c computes column c by right-looking substitution (same operation order as ms_potrf_inv)
            double x[2 * RP];
#pragma unroll
            for (int r = 0; r < 2 * RP; r++) x[r] = (r == lane) ? 1.0 : 0.0;
#pragma unroll
            for (int k = 0; k < 2 * RP; k++) {
                if (k >= d) break;
                const double xk = x[k] * rsb[k];
                x[k] = xk;
                const double2* col = reinterpret_cast<const double2*>(&Lb[k][0]);
#pragma unroll
                for (int r2 = (k + 1) / 2; r2 < RP; r2++) {
                    const double2 q = col[r2];
                    if (2 * r2 > k) x[2 * r2] -= q.x * xk;
                    x[2 * r2 + 1] -= q.y * xk;
                }
            }
            if (corner) {                            // D_N: the solve kernel reads inv(L_N) from the plain inverse storage
                if (lane < d) {
                    double* dst = Linv + m_offI[i] + (size_t)lane * d;
#pragma unroll
                    for (int r = 0; r < 2 * RP; r++) if (r < d) dst[r] = (r < lane) ? 0.0 : x[r];
                }
                continue;
            }
            const int D = m_cls[i];
            double* F = pk + m_pkF[i];
            double* Bk = pk + m_pkB[i];
            if (lane < d) {
#pragma unroll
                for (int r = 0; r < 2 * RP; r++) if (r < d) {
                    const double val = (r < lane) ? 0.0 : x[r];
                    F[r + lane * D] = val;           // inv(L)   column `lane`
                    Bk[lane + r * D] = val;          // inv(L)^T row `lane`
                }
            } else if (lane < d + o) {
                const int j = lane - d;
                const int D1 = m_cls[i + 1];
                double* Fn = pk + m_pkF[i + 1] + D1 * D1;          // B_i goes into the NEXT stage's forward packet [D1 x D]
                double* BT = Bk + D * D;                            // B_i^T [D x D1]
                for (int k = 0; k < d; k++) { const double val = Lb[k][lane]; Fn[j + k * D1] = val; BT[k + j * D] = val; }
            } else if (lane < rows) {
                const int f = lane - d - o;
                const int PD = i > 0 ? m_cls[i - 1] : 0, ND = (i + 2 < N) ? m_cls[i + 1] : 0;
                double* E = F + D * D + D * PD;                     // E_i [w x D]
                double* ET = Bk + D * D + D * ND;                   // E_i^T [D x WP]
                for (int k = 0; k < d; k++) { const double val = Lb[k][lane]; E[f + k * ww] = val; ET[k + f * D] = val; }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// factor without arrow (w == 0), the streamlined variant: the helper warp also PREPARES the fronts -- it copies the
// assembled K blocks of stage i+2 into a zero-padded 32x33 smem tile with cp.async while the chain warp works on stage
// i+1 -- and the chain warp stores its whole register row to a zero-extended 64x65 carry tile, so that the next
// stage's assembly is `v[c] = F[lane][c] + S[d_prev + lane][d_prev + c]` with no predicates at all.
// ------------------------------------------------------------------------------------------------------------------
struct MswChainSmem {
    double L[2][32][32];
    double rs[2][32];
    double F[2][32][33];
    double S[32][33];         // register rows of the previous stage; the next stage reads its Schur part at [d_prev + lane][d_prev + c]
};

__device__ __forceinline__ void msw_cp_async8(double* dst_smem, const double* src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(src) : "memory");
}

// Segment mode (parallel-in-horizon, multistage_partition.cuh): blockIdx.y selects a run of stages [seg_bounds[2y], seg_bounds[2y+1])
// that is factorised as an independent chain (the separator stages between the runs are left out); the Schur complement the last
// stage of the run leaves on its coupling rows is exported to carry_all[b][y][32 x 32] for the reduced (separator) system.
template <int RP>
__global__ void __launch_bounds__(64) msw_factor_chain_kernel(MsDev s, const double* __restrict__ fac_all, double* __restrict__ pk_all, size_t pk_stride,
                                                              const int* __restrict__ active, const int* __restrict__ seg_bounds, double* __restrict__ carry_all) {
    extern __shared__ __align__(16) unsigned char msw_raw[];
    MswChainSmem& sm = *reinterpret_cast<MswChainSmem*>(msw_raw);
    int* meta = reinterpret_cast<int*>(msw_raw + sizeof(MswChainSmem));
    const int b = blockIdx.x;
    if (active && !active[b]) return;
    const double* fac = fac_all + (size_t)b * s.total;
    double* pk = pk_all + (size_t)b * pk_stride;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int N = s.N;
    int i0 = 0, nst = N - 1;                          // stages [i0, nst)
    if (seg_bounds) { i0 = seg_bounds[2 * blockIdx.y]; nst = seg_bounds[2 * blockIdx.y + 1]; }
    // segment mode stages only the run's SLICE of the meta block (stages i0 .. nst, 12 x (len + 1) ints ~ 1 KB) so that five CTAs of
    // 43 KB fit one SM; a meta read from global memory would put an L2 round trip on every stage of the dependent chain
    int ms = N, mb = 0;                               // array stride and first stage of what `meta` holds
    if (!seg_bounds) { for (int e = threadIdx.x; e < MS_META * N; e += 64) meta[e] = s.start[e]; }
    else {
        ms = nst - i0 + 1; mb = i0;
        for (int e = threadIdx.x; e < MS_META * ms; e += 64) { const int a = e / ms, i = e - a * ms; meta[e] = s.start[a * N + mb + i]; }
    }
    for (int e = threadIdx.x; e < 32 * 33; e += 64) (&sm.S[0][0])[e] = 0.0;
    __syncthreads();
    const int *m_diag = meta + ms - mb, *m_off = meta + 2 * ms - mb, *m_offD = meta + 3 * ms - mb, *m_offB = meta + 4 * ms - mb, *m_cls = meta + 7 * ms - mb,
              *m_pkF = meta + 8 * ms - mb, *m_pkB = meta + 10 * ms - mb;
    // front of stage j -> F[j & 1] (one warp)
    auto prepare = [&](int j) {
        double* Fb = &sm.F[j & 1][0][0];
        for (int e = lane; e < 32 * 33; e += 32) Fb[e] = 0.0;
        __syncwarp();
        const int d = m_diag[j], o = (j + 2 < N) ? m_off[j] : 0;
        const double* src = nullptr; int ld = 0;
        if (lane < d) { src = fac + m_offD[j] + lane; ld = d; }
        else if (lane < d + o) { src = fac + m_offB[j] + (lane - d); ld = o; }
        if (src) for (int c = 0; c < d; c++) msw_cp_async8(Fb + lane * 33 + c, src + (size_t)c * ld);
        msw_cp_commit();
    };
    if (warp == 0) { if (nst > i0) prepare(i0); } else { if (nst > i0 + 1) prepare(i0 + 1); }
    msw_cp_wait<0>();
    __syncthreads();

    if (warp == 0) {
        int dprev = 0;
        for (int i = i0; i < nst; i++) {
            const int d = m_diag[i];
            double v[2 * RP];
            {
                const double* Frow = &sm.F[i & 1][lane][0];
                const bool rowok = dprev + lane < 32;
                const double* Srow = &sm.S[rowok ? dprev + lane : 0][dprev];
                const int cmax = rowok ? 32 - dprev : 0;          // the Schur complement of the previous front; zero beyond it
#pragma unroll
                for (int c = 0; c < 2 * RP; c++) v[c] = Frow[c] + (c < cmax ? Srow[c] : 0.0);
            }
            __syncwarp();
            double (*Lb)[32] = sm.L[i & 1];
            double* rsb = sm.rs[i & 1];
#pragma unroll
            for (int k = 0; k < 2 * RP; k++) {
                if (k >= d) break;
                const double piv = __shfl_sync(0xffffffffu, v[k], k);
                const double r = rsqrt(piv);
                const double l = v[k] * r;
                Lb[k][lane] = l;
                if (lane == 0) rsb[k] = r;
                __syncwarp();
                const double2* col = reinterpret_cast<const double2*>(&Lb[k][0]);
                double2 q[RP];
#pragma unroll
                for (int c2 = (k + 1) / 2; c2 < RP; c2++) q[c2] = col[c2];
#pragma unroll
                for (int c2 = (k + 1) / 2; c2 < RP; c2++) {
                    if (2 * c2 > k) v[2 * c2] -= l * q[c2].x;
                    v[2 * c2 + 1] -= l * q[c2].y;
                }
            }
            {
                double* Srow = &sm.S[lane][0];
#pragma unroll
                for (int c = 0; c < 2 * RP; c++) Srow[c] = v[c];
            }
            dprev = d;
            __syncthreads();
        }
        if (carry_all && nst > i0) {                 // Schur complement on the coupling rows of the run's last stage (lower part is meaningful)
            double* C = carry_all + ((size_t)b * gridDim.y + blockIdx.y) * 1024;
            for (int c = 0; c < 32; c++) C[lane + 32 * c] = (dprev + lane < 32 && dprev + c < 32) ? sm.S[dprev + lane][dprev + c] : 0.0;
        }
    } else {
        for (int i = i0; i < nst; i++) {
            __syncthreads();
            if (i + 2 < nst) prepare(i + 2);        // F[i & 1] was consumed by the chain warp at the start of stage i
            const int d = m_diag[i], o = (i + 2 < N) ? m_off[i] : 0;
            const double (*Lb)[32] = sm.L[i & 1];
            const double* rsb = sm.rs[i & 1];
            double x[2 * RP];
#pragma unroll
            for (int r = 0; r < 2 * RP; r++) x[r] = (r == lane) ? 1.0 : 0.0;
#pragma unroll
            for (int k = 0; k < 2 * RP; k++) {
                if (k >= d) break;
                const double xk = x[k] * rsb[k];
                x[k] = xk;
                const double2* col = reinterpret_cast<const double2*>(&Lb[k][0]);
#pragma unroll
                for (int r2 = (k + 1) / 2; r2 < RP; r2++) {
                    const double2 q = col[r2];
                    if (2 * r2 > k) x[2 * r2] -= q.x * xk;
                    x[2 * r2 + 1] -= q.y * xk;
                }
            }
            const int D = m_cls[i];
            double* F = pk + m_pkF[i];
            double* Bk = pk + m_pkB[i];
            if (lane < d) {
#pragma unroll
                for (int r = 0; r < 2 * RP; r++) if (r < d) {
                    const double val = (r < lane) ? 0.0 : x[r];
                    F[r + lane * D] = val;
                    Bk[lane + r * D] = val;
                }
            } else if (lane < d + o) {
                const int j = lane - d;
                const int D1 = m_cls[i + 1];
                double* Fn = pk + m_pkF[i + 1] + D1 * D1;
                double* BT = Bk + D * D;
                for (int k = 0; k < d; k++) { const double val = Lb[k][lane]; Fn[j + k * D1] = val; BT[k + j * D] = val; }
            }
            msw_cp_wait<0>();
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// solve: one warp per QP.  smem: xs[n + 32] | tmp[32] | accN[32] | meta | ring[R][slot]
// ------------------------------------------------------------------------------------------------------------------
constexpr int MSW_PF = 8;              // stages in flight
constexpr int MSW_R = MSW_PF;          // ring slots: a packet is self-contained, so its slot is free as soon as the stage is done

__device__ __forceinline__ void msw_issue(const double* src, int sz, double* slot, int lane) {
    for (int e = 2 * lane; e < sz; e += 64) msw_cp_async16(slot + e, src + e);
}
template <int D>
__device__ __forceinline__ double msw_reduce_h(double v) {
#pragma unroll
    for (int off = D; off < 32; off <<= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}
// acc = sum_k M[r + k*D] * y[k], k = h, h+H, ... < KD   (M column-major D x KD in smem, y in smem)
template <int D, int KD>
__device__ __forceinline__ double msw_matvec(const double* __restrict__ M, const double* __restrict__ y, int r, int h) {
    constexpr int H = 32 / D;
    double a0 = 0.0, a1 = 0.0;
#pragma unroll
    for (int kk = 0; kk < KD / H; kk += 2) {
        a0 += M[r + (kk * H + h) * D] * y[kk * H + h];
        if (kk + 1 < KD / H) a1 += M[r + ((kk + 1) * H + h) * D] * y[(kk + 1) * H + h];
    }
    return a0 + a1;
}
template <int D>
__device__ __forceinline__ double msw_matvec_dyn(const double* M, const double* y, int KD, int r, int h) {
    switch (KD) {
        case 8: return msw_matvec<D, 8>(M, y, r, h);
        case 16: return msw_matvec<D, 16>(M, y, r, h);
        case 32: return msw_matvec<D, 32>(M, y, r, h);
        default: return 0.0;
    }
}

template <int D>
__device__ __forceinline__ void msw_fwd_stage(const double* pkt, int PD, double* xs, int st, int pst, int d, int w, double* tmp, double* accN, int lane) {
    constexpr int H = 32 / D;
    const int r = lane % D, h = lane / D;
    double acc = msw_matvec_dyn<D>(pkt + D * D, xs + pst, PD, r, h);            // B_{i-1} y_{i-1}
    acc = msw_reduce_h<D>(acc);
    if (h == 0) tmp[r] = xs[st + r] - acc;                                       // rows >= d: finite junk, hits zero columns
    __syncwarp();
    acc = msw_matvec<D, D>(pkt, tmp, r, h);                                      // inv(L_i) v
    acc = msw_reduce_h<D>(acc);
    if (h == 0 && r < d) xs[st + r] = acc;
    __syncwarp();
    if (w > 0) {                                                                 // arrow: x_N -= E_i y_i (off the chain)
        if (lane < w) {
            const double* E = pkt + D * D + D * PD;
            double a = 0.0;
            for (int k = 0; k < d; k++) a += E[lane + k * w] * xs[st + k];
            accN[lane] += a;
        }
    }
    (void)H;
}
template <int D>
__device__ __forceinline__ void msw_bwd_stage(const double* pkt, int ND, double* xs, int st, int nst, int d, int w, int n, double* tmp, int lane) {
    const int r = lane % D, h = lane / D;
    double acc = msw_matvec_dyn<D>(pkt + D * D, xs + nst, ND, r, h);            // B_i^T x_{i+1}
    if (w > 0) {
        const double* ET = pkt + D * D + D * ND;
        if (h == 0) for (int f = 0; f < w; f++) acc += ET[r + f * D] * xs[n - w + f];
    }
    acc = msw_reduce_h<D>(acc);
    if (h == 0) tmp[r] = (r < d) ? xs[st + r] - acc : 0.0;
    __syncwarp();
    acc = msw_matvec<D, D>(pkt, tmp, r, h);                                      // inv(L_i)^T v
    acc = msw_reduce_h<D>(acc);
    if (h == 0 && r < d) xs[st + r] = acc;
    __syncwarp();
}

// one warp, one QP: x <- L^-T L^-1 x.  `xs` is this warp's shared-memory workspace (layout above); also called by the fused
// parallel-in-horizon solve (multistage_partition.cuh) for the reduced chain.
__device__ __forceinline__ void msw_solve_body(const MsDev& s, int slot_doubles, const double* __restrict__ Linv, const double* __restrict__ pk,
                                               double* __restrict__ x, double* xs, int lane) {
    const int N = s.N, w = s.w, n = s.n;
    const int NS = N - 1;                          // regular stages
    double* tmp = xs + ((n + 1) & ~1) + 32;        // 32 doubles of slack after xs: padded rows/columns read (and ignore) them
    double* accN = tmp + 32;
    int* meta = reinterpret_cast<int*>(accN + 32);
    double* ring = accN + 32 + ((MS_META * N + 1) / 2 + 1) / 2 * 2;
    for (int e = lane; e < MS_META * N; e += 32) meta[e] = s.start[e];
    const int *m_start = meta, *m_diag = meta + N, *m_cls = meta + 7 * N, *m_pkF = meta + 8 * N, *m_szF = meta + 9 * N, *m_pkB = meta + 10 * N, *m_szB = meta + 11 * N;
    __syncwarp();

    // ---------------- forward ----------------
    for (int q = 0; q < MSW_PF; q++) { if (q < NS) msw_issue(pk + m_pkF[q], m_szF[q], ring + (size_t)(q % MSW_R) * slot_doubles, lane); msw_cp_commit(); }
    for (int i = lane; i < n; i += 32) xs[i] = x[i];
    for (int i = n + lane; i < ((n + 1) & ~1) + 32; i += 32) xs[i] = 0.0;
    if (lane < 32) accN[lane] = 0.0;
    for (int i = 0; i < NS; i++) {
        msw_cp_wait<MSW_PF - 1>();
        __syncwarp();
        const int d = m_diag[i], st = m_start[i], D = m_cls[i];
        const int PD = i > 0 ? m_cls[i - 1] : 0, pst = i > 0 ? m_start[i - 1] : 0;
        const double* pkt = ring + (size_t)(i % MSW_R) * slot_doubles;
        if (D == 16) msw_fwd_stage<16>(pkt, PD, xs, st, pst, d, w, tmp, accN, lane);
        else if (D == 8) msw_fwd_stage<8>(pkt, PD, xs, st, pst, d, w, tmp, accN, lane);
        else msw_fwd_stage<32>(pkt, PD, xs, st, pst, d, w, tmp, accN, lane);
        __syncwarp();
        const int j = i + MSW_PF;
        if (j < NS) msw_issue(pk + m_pkF[j], m_szF[j], ring + (size_t)(j % MSW_R) * slot_doubles, lane);
        msw_cp_commit();
    }
    msw_cp_wait<0>();
    __syncwarp();
    // ---------------- arrow corner ----------------
    if (w > 0) {
        const double* I = Linv + s.offI[N - 1];
        if (lane < w) tmp[lane] = xs[n - w + lane] - accN[lane];
        __syncwarp();
        double y = 0.0;
        if (lane < w) for (int k = 0; k <= lane; k++) y += I[lane + k * w] * tmp[k];
        __syncwarp();
        if (lane < w) tmp[lane] = y;
        __syncwarp();
        if (lane < w) { double a = 0.0; for (int k = lane; k < w; k++) a += I[k + lane * w] * tmp[k]; xs[n - w + lane] = a; }
        __syncwarp();
    }
    // ---------------- backward ----------------
    for (int q = 0; q < MSW_PF; q++) { const int i = NS - 1 - q; if (i >= 0) msw_issue(pk + m_pkB[i], m_szB[i], ring + (size_t)(i % MSW_R) * slot_doubles, lane); msw_cp_commit(); }
    for (int i = NS - 1; i >= 0; i--) {
        msw_cp_wait<MSW_PF - 1>();
        __syncwarp();
        const int d = m_diag[i], st = m_start[i], D = m_cls[i];
        const int ND = (i + 2 < N) ? m_cls[i + 1] : 0, nst = (i + 2 < N) ? m_start[i + 1] : 0;
        const double* pkt = ring + (size_t)(i % MSW_R) * slot_doubles;
        if (D == 16) msw_bwd_stage<16>(pkt, ND, xs, st, nst, d, w, n, tmp, lane);
        else if (D == 8) msw_bwd_stage<8>(pkt, ND, xs, st, nst, d, w, n, tmp, lane);
        else msw_bwd_stage<32>(pkt, ND, xs, st, nst, d, w, n, tmp, lane);
        const int j = i - MSW_PF;
        if (j >= 0) msw_issue(pk + m_pkB[j], m_szB[j], ring + (size_t)(j % MSW_R) * slot_doubles, lane);
        msw_cp_commit();
    }
    msw_cp_wait<0>();
    __syncwarp();
    for (int i = lane; i < n; i += 32) x[i] = xs[i];
}

__global__ void __launch_bounds__(32) msw_solve_kernel(MsDev s, int slot_doubles, const double* __restrict__ inv_all, const double* __restrict__ pk_all,
                                                       size_t pk_stride, double* __restrict__ X, const int* __restrict__ active) {
    extern __shared__ __align__(16) double msw_solve_smem[];
    const int b = blockIdx.x;
    if (active && !active[b]) return;
    msw_solve_body(s, slot_doubles, inv_all ? inv_all + (size_t)b * s.total_inv : nullptr, pk_all + (size_t)b * pk_stride, X + (size_t)b * s.n, msw_solve_smem, threadIdx.x);
}

}  // namespace b200
