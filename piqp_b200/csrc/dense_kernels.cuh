// piqp_b200/csrc/dense_kernels.cuh -- batched dense fp64 kernels for the KKT hot path (sm_100a).
//
// What they replace in the reference (include/piqp/dense/kkt.hpp):
//   gemm_nt_tile_kernel<EPI_ASSEMBLE>  dense::KKT::update_kkt            :140-160  (K = P + diag + AtA/delta + G^T Z^-1 G)
//   gemm_nt_tile_kernel<EPI_STORE>     AT_A = AT * AT^T                  :53,68
//   gemm_nt_tile_kernel<EPI_SUB> + potf2_kernel + trsm_kernel            Eigen::LLT::compute, :82 (blocked Cholesky)
//   trsv_kernel                         llt.solveInPlace                  :170
//   gemv_n_kernel / gemv_t_kernel       the GT / AT products of solve() and eval_*()  :92-104, :108-132
//
// FP64 on Blackwell: tcgen05.mma has no f64 kind (ptxas rejects .kind::f64), so the tensor-core path for
// double precision is the warp-level DMMA (mma.sync.aligned.m8n8k4.f64).  The contraction kernels below
// stage 128 x 16 operand panels through shared memory with a 3-stage cp.async pipeline and issue DMMA
// from 8 warps (2 x 4 warp grid, 64 x 32 warp tile, 64 fp64 accumulators per lane).
//
// Layout: every matrix is column-major with a padded leading dimension ld (multiple of 8 doubles); the
// padding rows are zero.  Batched arrays are instance-major with a fixed stride.
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace b200 {

constexpr int TILE = 128;       // C tile is TILE x TILE
constexpr int KB = 16;          // k-depth of one pipeline stage
constexpr int STAGES = 3;
constexpr int LDS_T = TILE + 4; // padded smem row (doubles): conflict-free DMMA fragment loads
constexpr int GEMM_THREADS = 256;
constexpr size_t GEMM_STAGE_SMEM = (size_t)STAGES * 2 * KB * LDS_T * sizeof(double);
constexpr int TS_LD = TILE + 4; // leading dimension of a full 128 x 128 tile staged in smem: Ts[col * TS_LD + row]
constexpr size_t TILE_SMEM = (size_t)TILE * TS_LD * sizeof(double);
constexpr size_t GEMM_SMEM = TILE_SMEM > GEMM_STAGE_SMEM ? TILE_SMEM : GEMM_STAGE_SMEM;

enum Epilogue { EPI_STORE = 0, EPI_SUB = 1, EPI_ASSEMBLE = 2 };

struct GemmArgs {
    // operands: acc(i,j) = sum_{k<K} A[(rowA0+i) + k*lda] * w[k] * B[(rowB0+j) + k*ldb]
    const double* A; long long strideA; int lda;
    const double* B; long long strideB; int ldb;
    const double* w; long long stridew;   // nullable (no scaling)
    double* C; long long strideC; int ldc;
    int n;            // logical dimension of C (rows/cols < n are stored)
    int rows_valid;   // rows of A/B that exist in memory (>= n, padded ld); loads beyond are zero-filled
    int K;            // contraction length
    int nt;           // tiles per dimension
    int tj_fixed;     // >= 0: column mode (tiles (tj_fixed + t, tj_fixed)); < 0: all lower tiles of the tile columns >= tj_start
    int tj_start;
    int ncol;         // > 0: only columns < ncol of C are stored (narrow window updates); 0: all n
    int tiles;        // tiles per instance
    int t0;           // gemm_nt_t64_kernel: index of the first tile of this launch within the enumeration (skip the diagonal tile: t0 = 2)
    // EPI_ASSEMBLE extras
    const double* Pf; long long strideP;     // full symmetric P, ld = ldc
    const double* AtA; long long strideAtA;  // nullable
    const double* xreg; long long stridex;
    const double* delta;                      // [batch]
    const int* active;                        // nullable per-instance mask
    const int* fail;                          // nullable: skip instances whose factorisation already failed
};

#ifdef __CUDACC__

__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem, int src_bytes) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(sa), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// stage one TILE x KB panel: smem[k][r] <- M[(row0 + r) + (k0 + k) * ld], zero-filled outside [0,rows_valid) x [0,K)
__device__ __forceinline__ void load_panel(double* sm, const double* M, int ld, int row0, int rows_valid, int k0, int K) {
    // TILE/2 = 64 16-byte chunks per k-row, KB rows -> 1024 chunks, 4 per thread
#pragma unroll
    for (int it = 0; it < (KB * TILE / 2) / GEMM_THREADS; it++) {
        const int c = threadIdx.x + it * GEMM_THREADS;
        const int k = c / (TILE / 2);
        const int r = (c % (TILE / 2)) * 2;
        const int gr = row0 + r, gk = k0 + k;
        const bool ok = (gr < rows_valid) && (gk < K);
        const double* src = ok ? (M + (size_t)gk * ld + gr) : M;
        cp_async16(sm + k * LDS_T + r, src, ok ? 16 : 0);
    }
}

// acc(i,j) += sum_{k<K} A[(rowA0+i) + k*lda] * w[k] * B[(rowB0+j) + k*ldb] for the 128 x 128 tile; warp grid 2 (M) x 4 (N),
// warp tile 64 x 32, lane holds C(row = gq, cols 2*tq, 2*tq+1) of each 8x8 fragment.  Ends with all cp.async drained.
template <bool HAS_W>
__device__ __forceinline__ void gemm_mainloop(double (&acc)[8][4][2], double* smem, const double* A, int lda, int rowA0,
                                              const double* B, int ldb, int rowB0, int rows_valid, int K, const double* w, bool diag) {
    double* As = smem;                              // [STAGES][KB][LDS_T]
    double* Bs = smem + STAGES * KB * LDS_T;        // [STAGES][KB][LDS_T]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wm = warp >> 2, wn = warp & 3;
    const int gq = lane >> 2, tq = lane & 3;
    const int nkb = (K + KB - 1) / KB;
#pragma unroll
    for (int s = 0; s < STAGES - 1; s++) {
        if (s < nkb) {
            load_panel(As + s * KB * LDS_T, A, lda, rowA0, rows_valid, s * KB, K);
            if (!diag) load_panel(Bs + s * KB * LDS_T, B, ldb, rowB0, rows_valid, s * KB, K);
        }
        cp_async_commit();
    }
    // per-lane scale factors w[k] for the k rows this lane touches (k = kb*KB + kk*4 + tq), fetched one k-block ahead
    double wcur[KB / 4], wnext[KB / 4];
#pragma unroll
    for (int kk = 0; kk < KB / 4; kk++) { wcur[kk] = 1.0; wnext[kk] = 1.0; }
    if (HAS_W) {
#pragma unroll
        for (int kk = 0; kk < KB / 4; kk++) { const int gk = kk * 4 + tq; wcur[kk] = gk < K ? w[gk] : 0.0; }
    }
    for (int kb = 0; kb < nkb; kb++) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            const int nx = kb + STAGES - 1;
            if (nx < nkb) {
                const int s = nx % STAGES;
                load_panel(As + s * KB * LDS_T, A, lda, rowA0, rows_valid, nx * KB, K);
                if (!diag) load_panel(Bs + s * KB * LDS_T, B, ldb, rowB0, rows_valid, nx * KB, K);
            }
            cp_async_commit();
        }
        if (HAS_W) {
#pragma unroll
            for (int kk = 0; kk < KB / 4; kk++) { const int gk = (kb + 1) * KB + kk * 4 + tq; wnext[kk] = gk < K ? w[gk] : 0.0; }
        }
        const int s = kb % STAGES;
        const double* as = As + s * KB * LDS_T;
        const double* bs = diag ? as : (Bs + s * KB * LDS_T);
#pragma unroll
        for (int kk = 0; kk < KB / 4; kk++) {
            double af[8], bf[4];
            const int krow = kk * 4 + tq;
#pragma unroll
            for (int i = 0; i < 8; i++) af[i] = as[krow * LDS_T + wm * 64 + i * 8 + gq];
#pragma unroll
            for (int j = 0; j < 4; j++) { bf[j] = bs[krow * LDS_T + wn * 32 + j * 8 + gq]; if (HAS_W) bf[j] *= wcur[kk]; }
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
        if (HAS_W) {
#pragma unroll
            for (int kk = 0; kk < KB / 4; kk++) wcur[kk] = wnext[kk];
        }
    }
    cp_async_wait<0>();
}

// stage the accumulators of the 2x4 warp grid into a full tile in smem: Cs[col * TS_LD + row] = sign * acc
__device__ __forceinline__ void acc_to_smem(const double (&acc)[8][4][2], double* Cs, double sign) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wm = warp >> 2, wn = warp & 3, gq = lane >> 2, tq = lane & 3;
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int e = 0; e < 2; e++)
                Cs[(wn * 32 + j * 8 + tq * 2 + e) * TS_LD + wm * 64 + i * 8 + gq] = sign * acc[i][j][e];
}

template <int EPI, bool HAS_W>
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_nt_tile_kernel(GemmArgs g) {
    extern __shared__ __align__(16) double smem[];
    const int b = blockIdx.x / g.tiles;
    int t = blockIdx.x % g.tiles;
    if (g.active && !g.active[b]) return;
    if (g.fail && g.fail[b]) return;
    int ti, tj;
    if (g.tj_fixed >= 0) { tj = g.tj_fixed; ti = tj + t; }
    else { tj = g.tj_start; while (t >= g.nt - tj) { t -= g.nt - tj; tj++; } ti = tj + t; }
    const bool diag = (ti == tj) && (g.A == g.B);
    const double* A = g.A + (size_t)b * g.strideA;
    const double* B = g.B + (size_t)b * g.strideB;
    const double* w = HAS_W ? g.w + (size_t)b * g.stridew : nullptr;
    const int rowA0 = ti * TILE, rowB0 = tj * TILE;

    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
    gemm_mainloop<HAS_W>(acc, smem, A, g.lda, rowA0, B, g.ldb, rowB0, g.rows_valid, g.K, w, diag);

    // ---- epilogue through shared memory: coalesced 16-byte global accesses, loads batched ahead of the stores
    __syncthreads();
    acc_to_smem(acc, smem, 1.0);
    __syncthreads();
    double* __restrict__ C = g.C + (size_t)b * g.strideC;
    const double dinv = (EPI == EPI_ASSEMBLE) ? 1.0 / g.delta[b] : 0.0;
    const double* __restrict__ Pf = (EPI == EPI_ASSEMBLE) ? g.Pf + (size_t)b * g.strideP : nullptr;
    const double* __restrict__ AtA = (EPI == EPI_ASSEMBLE && g.AtA) ? g.AtA + (size_t)b * g.strideAtA : nullptr;
    const double* __restrict__ xr = (EPI == EPI_ASSEMBLE) ? g.xreg + (size_t)b * g.stridex : nullptr;
    const int r = rowA0 + (threadIdx.x & 63) * 2;
    constexpr int U = 4;
#pragma unroll 1
    for (int it0 = 0; it0 < TILE / 4; it0 += U) {
        double2 base[U], extra[U];
        bool ok[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int cl = (it0 + u) * 4 + (threadIdx.x >> 6);
            const int c = rowB0 + cl;
            ok[u] = (r + 1 < g.rows_valid + 0) && (c < (g.ncol > 0 ? g.ncol : g.n)) && (r < g.n) && (r + 1 >= c);
            base[u] = make_double2(0.0, 0.0); extra[u] = make_double2(0.0, 0.0);
            if (ok[u]) {
                const size_t idx = (size_t)c * g.ldc + r;
                if (EPI == EPI_SUB) base[u] = *reinterpret_cast<const double2*>(C + idx);
                if (EPI == EPI_ASSEMBLE) { base[u] = *reinterpret_cast<const double2*>(Pf + idx); if (AtA) extra[u] = *reinterpret_cast<const double2*>(AtA + idx); }
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (!ok[u]) continue;
            const int cl = (it0 + u) * 4 + (threadIdx.x >> 6);
            const int c = rowB0 + cl;
            const double2 a = *reinterpret_cast<const double2*>(smem + cl * TS_LD + (threadIdx.x & 63) * 2);
            double v0 = a.x, v1 = a.y;
            if (EPI == EPI_SUB) { v0 = base[u].x - v0; v1 = base[u].y - v1; }
            if (EPI == EPI_ASSEMBLE) {
                double b0 = base[u].x, b1 = base[u].y;
                if (r == c) b0 += xr[r];
                if (r + 1 == c) b1 += xr[r + 1];
                if (AtA) { b0 += dinv * extra[u].x; b1 += dinv * extra[u].y; }
                v0 = b0 + v0; v1 = b1 + v1;
            }
            const size_t idx = (size_t)c * g.ldc + r;
            if (r >= c && r + 1 < g.n) *reinterpret_cast<double2*>(C + idx) = make_double2(v0, v1);
            else { if (r >= c && r < g.n) C[idx] = v0; if (r + 1 >= c && r + 1 < g.n) C[idx + 1] = v1; }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// gemm_nt_t64_kernel: the same contraction with a 128 (rows) x 64 (columns) C tile, 8 warps as 4 (M) x 2 (N) with 32 x 32
// warp tiles -> 32 fp64 accumulators per lane, <= 128 registers, 77 KB of shared memory: TWO CTAs PER SM.  The 128 x 128 kernel
// above holds one CTA per SM, so its prologue (pipeline fill), its barriers and its epilogue (read-modify-write of the C tile)
// leave the DMMA pipe idle (ncu: 65 % tensor-pipe active); with two co-resident CTAs one tile's epilogue overlaps the other's
// main loop.  Tiles: per 128-column tile column tj (units of the 128-kernel, so tj_start / tiles / ncol keep their meaning
// for the callers) the rows ti >= tj, each split into two 64-column halves.
// ---------------------------------------------------------------------------------------------------
constexpr int T64_N = 64;
constexpr int LDS_B64 = T64_N + 4;
constexpr size_t T64_STAGE_SMEM = (size_t)STAGES * KB * (LDS_T + LDS_B64 + 1) * sizeof(double);     // + the stage's KB weights
constexpr size_t T64_TILE_SMEM = (size_t)T64_N * TS_LD * sizeof(double);
constexpr size_t T64_SMEM = T64_TILE_SMEM > T64_STAGE_SMEM ? T64_TILE_SMEM : T64_STAGE_SMEM;

// stage a 64 x KB panel: smem[k][r] <- M[(row0 + r) + (k0 + k) * ld]
__device__ __forceinline__ void load_panel64(double* sm, const double* M, int ld, int row0, int rows_valid, int k0, int K) {
#pragma unroll
    for (int it = 0; it < (KB * T64_N / 2) / GEMM_THREADS; it++) {
        const int c = threadIdx.x + it * GEMM_THREADS;
        const int k = c / (T64_N / 2);
        const int r = (c % (T64_N / 2)) * 2;
        const int gr = row0 + r, gk = k0 + k;
        const bool ok = (gr < rows_valid) && (gk < K);
        const double* src = ok ? (M + (size_t)gk * ld + gr) : M;
        cp_async16(sm + k * LDS_B64 + r, src, ok ? 16 : 0);
    }
}

template <int EPI, bool HAS_W>
__global__ void __launch_bounds__(GEMM_THREADS, 2) gemm_nt_t64_kernel(GemmArgs g) {
    extern __shared__ __align__(16) double smem[];
    const int b = blockIdx.x / g.tiles;
    int t = blockIdx.x % g.tiles + g.t0;
    if (g.active && !g.active[b]) return;
    if (g.fail && g.fail[b]) return;
    int tj = g.tj_start;
    while (t >= 2 * (g.nt - tj)) { t -= 2 * (g.nt - tj); tj++; }
    const int ti = tj + (t >> 1), half = t & 1;
    const bool diag = (ti == tj) && (g.A == g.B);          // the B rows are a 64-row slice of the A panel: read it from there
    const double* A = g.A + (size_t)b * g.strideA;
    const double* B = g.B + (size_t)b * g.strideB;
    const double* w = HAS_W ? g.w + (size_t)b * g.stridew : nullptr;
    const int rowA0 = ti * TILE, rowB0 = tj * TILE + half * T64_N;
    const int K = g.K, rows_valid = g.rows_valid;

    double* As = smem;                                   // [STAGES][KB][LDS_T]
    double* Bs = smem + STAGES * KB * LDS_T;             // [STAGES][KB][LDS_B64]
    double* Ws = Bs + STAGES * KB * LDS_B64;             // [STAGES][KB]: the weights of the stage ride the same cp.async groups (a global load per
                                                         // k-block sat on the DMUL -> DMMA chain: ncu source view, 4 % of the stall samples in long scoreboard)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wm = warp >> 1, wn = warp & 1;
    const int gq = lane >> 2, tq = lane & 3;
    const int nkb = (K + KB - 1) / KB;
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
#pragma unroll
    for (int s = 0; s < STAGES - 1; s++) {
        if (s < nkb) {
            load_panel(As + s * KB * LDS_T, A, g.lda, rowA0, rows_valid, s * KB, K);
            if (!diag) load_panel64(Bs + s * KB * LDS_B64, B, g.ldb, rowB0, rows_valid, s * KB, K);
            if (HAS_W && threadIdx.x < KB) { const int gk = s * KB + threadIdx.x; cp_async8(Ws + s * KB + threadIdx.x, gk < K ? w + gk : w, gk < K ? 8 : 0); }
        }
        cp_async_commit();
    }
    const int ldsb = diag ? LDS_T : LDS_B64;
    // warp tiles of a diagonal block that lie strictly above the diagonal are never stored: their warps skip the DMMAs (they
    // still load and synchronise) and leave the tensor pipe to the co-resident CTA
    const bool skip_mma = (ti == tj) && (half * T64_N + wn * 32 >= wm * 32 + 32);
    for (int kb = 0; kb < nkb; kb++) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            const int nx = kb + STAGES - 1;
            if (nx < nkb) {
                const int s = nx % STAGES;
                load_panel(As + s * KB * LDS_T, A, g.lda, rowA0, rows_valid, nx * KB, K);
                if (!diag) load_panel64(Bs + s * KB * LDS_B64, B, g.ldb, rowB0, rows_valid, nx * KB, K);
                if (HAS_W && threadIdx.x < KB) { const int gk = nx * KB + threadIdx.x; cp_async8(Ws + s * KB + threadIdx.x, gk < K ? w + gk : w, gk < K ? 8 : 0); }
            }
            cp_async_commit();
        }
        const int s = kb % STAGES;
        double wcur[KB / 4];
        if (HAS_W) {
#pragma unroll
            for (int kk = 0; kk < KB / 4; kk++) wcur[kk] = Ws[s * KB + kk * 4 + tq];
        }
        const double* as = As + s * KB * LDS_T;
        const double* bs = diag ? (as + half * T64_N) : (Bs + s * KB * LDS_B64);
        if (!skip_mma) {
#pragma unroll
        for (int kk = 0; kk < KB / 4; kk++) {
            double af[4], bf[4];
            const int krow = kk * 4 + tq;
#pragma unroll
            for (int i = 0; i < 4; i++) af[i] = as[krow * LDS_T + wm * 32 + i * 8 + gq];
#pragma unroll
            for (int j = 0; j < 4; j++) { bf[j] = bs[krow * ldsb + wn * 32 + j * 8 + gq]; if (HAS_W) bf[j] *= wcur[kk]; }
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
        }
    }
    cp_async_wait<0>();
    // ---- epilogue through shared memory (Cs[col * TS_LD + row], 64 columns): coalesced 16-byte global accesses
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int e = 0; e < 2; e++) smem[(wn * 32 + j * 8 + tq * 2 + e) * TS_LD + wm * 32 + i * 8 + gq] = acc[i][j][e];
    __syncthreads();
    double* __restrict__ C = g.C + (size_t)b * g.strideC;
    const double dinv = (EPI == EPI_ASSEMBLE) ? 1.0 / g.delta[b] : 0.0;
    const double* __restrict__ Pf = (EPI == EPI_ASSEMBLE) ? g.Pf + (size_t)b * g.strideP : nullptr;
    const double* __restrict__ AtA = (EPI == EPI_ASSEMBLE && g.AtA) ? g.AtA + (size_t)b * g.strideAtA : nullptr;
    const double* __restrict__ xr = (EPI == EPI_ASSEMBLE) ? g.xreg + (size_t)b * g.stridex : nullptr;
    const int ncol = g.ncol > 0 ? g.ncol : g.n;
    const int r = rowA0 + (threadIdx.x & 63) * 2;
    constexpr int U = 4;
#pragma unroll 1
    for (int it0 = 0; it0 < T64_N / 4; it0 += U) {
        double2 base[U], extra[U];
        bool ok[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int cl = (it0 + u) * 4 + (threadIdx.x >> 6);
            const int c = rowB0 + cl;
            ok[u] = (r + 1 < g.rows_valid) && (c < ncol) && (r < g.n) && (r + 1 >= c);
            base[u] = make_double2(0.0, 0.0); extra[u] = make_double2(0.0, 0.0);
            if (ok[u]) {
                const size_t idx = (size_t)c * g.ldc + r;
                if (EPI == EPI_SUB) base[u] = *reinterpret_cast<const double2*>(C + idx);
                if (EPI == EPI_ASSEMBLE) { base[u] = *reinterpret_cast<const double2*>(Pf + idx); if (AtA) extra[u] = *reinterpret_cast<const double2*>(AtA + idx); }
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (!ok[u]) continue;
            const int cl = (it0 + u) * 4 + (threadIdx.x >> 6);
            const int c = rowB0 + cl;
            const double2 a = *reinterpret_cast<const double2*>(smem + cl * TS_LD + (threadIdx.x & 63) * 2);
            double v0 = a.x, v1 = a.y;
            if (EPI == EPI_SUB) { v0 = base[u].x - v0; v1 = base[u].y - v1; }
            if (EPI == EPI_ASSEMBLE) {
                double b0 = base[u].x, b1 = base[u].y;
                if (r == c) b0 += xr[r];
                if (r + 1 == c) b1 += xr[r + 1];
                if (AtA) { b0 += dinv * extra[u].x; b1 += dinv * extra[u].y; }
                v0 = b0 + v0; v1 = b1 + v1;
            }
            const size_t idx = (size_t)c * g.ldc + r;
            if (r >= c && r + 1 < g.n) *reinterpret_cast<double2*>(C + idx) = make_double2(v0, v1);
            else { if (r >= c && r < g.n) C[idx] = v0; if (r + 1 >= c && r + 1 < g.n) C[idx + 1] = v1; }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// gemm_nt_t64_bulk_kernel: the same 128 x 64 tile kernel with the operand panels staged by the TMA engine's bulk copies
// (cp.async.bulk.shared.global, one 1 KB / 512 B row of the panel per copy: 32 copies + the 16 weights per stage) into a 4-stage ring
// guarded by mbarriers: full[s] (expect_tx = the stage's bytes) releases the consumers, empty[s] (one arrival per consumer warp)
// releases the slot.  A NINTH, PRODUCER-ONLY WARP issues the copies and runs the whole ring ahead of the eight consumer warps
// (288 threads at 96 registers: still two CTAs per SM); with BULK_PRODUCER_THREADS = 0 warp 0 doubles as the producer and refills a
// slot two k-blocks after its consumption (measured: assembly 68.6 instead of 62.6 ms per step of config 2).
// There is NO CTA-wide barrier in the main loop: the source view of the cp.async kernel (profiles/r02d_ncu_source_stalls.txt) had
// 13 % of its stall samples on the per-k-block __syncthreads, because eight warps that drift apart on the DMMA pipe were re-aligned
// 32 times per tile.  The padded rows of the cp.async layout are kept (conflict-free fragment loads): that is why the copies are 1-D
// bulk copies per panel row and not one tensor-map box per panel.  Rows of a partial tile that do not exist in memory are not copied;
// what the slot holds there only feeds accumulator rows / columns that the epilogue never stores.
// Preconditions (host-checked, else the cp.async kernel runs): K % 16 == 0, an even number of rows in memory, 16-byte aligned bases / strides.
// ---------------------------------------------------------------------------------------------------
constexpr int BULK_STAGES = 4;
constexpr int BULK_PRODUCER_THREADS = 32;      // 32: a ninth, producer-only warp; 0: warp 0 doubles as the producer
constexpr int BULK_STAGE_DOUBLES = KB * (LDS_T + LDS_B64) + KB;          // A panel | B panel | weights
constexpr size_t T64_BULK_SMEM_RING = (size_t)BULK_STAGES * BULK_STAGE_DOUBLES * sizeof(double);
constexpr size_t T64_BULK_SMEM = (T64_TILE_SMEM > T64_BULK_SMEM_RING ? T64_TILE_SMEM : T64_BULK_SMEM_RING) + 2 * BULK_STAGES * sizeof(unsigned long long);

__device__ __forceinline__ unsigned bk_smem(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bk_mbar_init(unsigned long long* bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bk_smem(bar)), "r"(count)); }
__device__ __forceinline__ void bk_mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bk_smem(bar)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bk_expect_tx(unsigned long long* bar, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bk_smem(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void bk_arrive(unsigned long long* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bk_smem(bar)) : "memory"); }
__device__ __forceinline__ void bk_bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(bk_smem(dst)), "l"(src), "r"(bytes), "r"(bk_smem(bar)) : "memory");
}

template <int EPI, bool HAS_W>
__global__ void __launch_bounds__(GEMM_THREADS + BULK_PRODUCER_THREADS, 2) gemm_nt_t64_bulk_kernel(GemmArgs g) {
    extern __shared__ __align__(128) double smem[];
    const int b = blockIdx.x / g.tiles;
    int t = blockIdx.x % g.tiles + g.t0;
    if (g.active && !g.active[b]) return;
    if (g.fail && g.fail[b]) return;
    int tj = g.tj_start;
    while (t >= 2 * (g.nt - tj)) { t -= 2 * (g.nt - tj); tj++; }
    const int ti = tj + (t >> 1), half = t & 1;
    const bool diag = (ti == tj) && (g.A == g.B);          // the B rows are a 64-row slice of the A panel: read it from there
    const double* A = g.A + (size_t)b * g.strideA;
    const double* B = g.B + (size_t)b * g.strideB;
    const double* w = HAS_W ? g.w + (size_t)b * g.stridew : nullptr;
    const int rowA0 = ti * TILE, rowB0 = tj * TILE + half * T64_N;
    const int K = g.K;
    unsigned long long* full = reinterpret_cast<unsigned long long*>(reinterpret_cast<unsigned char*>(smem) + (T64_BULK_SMEM - 2 * BULK_STAGES * sizeof(unsigned long long)));
    unsigned long long* empty = full + BULK_STAGES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wm = warp >> 1, wn = warp & 1;
    const int gq = lane >> 2, tq = lane & 3;
    const int nkb = K / KB;
    if (threadIdx.x == 0) {
        for (int s = 0; s < BULK_STAGES; s++) { bk_mbar_init(&full[s], 1); bk_mbar_init(&empty[s], GEMM_THREADS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // rows of the tile that exist in memory (even counts: 16-byte copies); the rows behind them keep whatever the slot held -- they only feed
    // accumulator rows / columns >= n, which the epilogue never stores
    const int ra = min(TILE, g.rows_valid - rowA0), rb = diag ? 0 : max(0, min(T64_N, g.rows_valid - rowB0));
    const unsigned stage_bytes = (unsigned)(sizeof(double) * KB * (ra + rb + (HAS_W ? 1 : 0)));
    // fill slot (kb % BULK_STAGES) with k-block kb: lanes 0..15 copy the A rows, 16..31 the B rows, lane 0 also the 16 weights
    auto fill = [&](int kb) {
        const int s = kb % BULK_STAGES;
        double* as = smem + (size_t)s * BULK_STAGE_DOUBLES;
        double* bs = as + KB * LDS_T;
        double* ws = bs + KB * LDS_B64;
        if (lane == 0) bk_expect_tx(&full[s], stage_bytes);
        __syncwarp();
        const int k = lane & 15;
        if (lane < 16) bk_bulk_g2s(as + k * LDS_T, A + (size_t)(kb * KB + k) * g.lda + rowA0, ra * sizeof(double), &full[s]);
        else if (rb > 0) bk_bulk_g2s(bs + k * LDS_B64, B + (size_t)(kb * KB + k) * g.ldb + rowB0, rb * sizeof(double), &full[s]);
        if (HAS_W && lane == 0) bk_bulk_g2s(ws, w + kb * KB, KB * sizeof(double), &full[s]);
    };
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
    if (BULK_PRODUCER_THREADS > 0 && warp == GEMM_THREADS / 32) {      // dedicated producer warp: runs ahead of the consumers by the whole ring
        for (int kb = 0; kb < nkb; kb++) {
            if (kb >= BULK_STAGES) bk_mbar_wait(&empty[kb % BULK_STAGES], ((kb / BULK_STAGES) - 1) & 1);
            fill(kb);
        }
    }
    if (BULK_PRODUCER_THREADS == 0 && warp == 0) { for (int kb = 0; kb < BULK_STAGES - 2 && kb < nkb; kb++) fill(kb); }
    const int ldsb = diag ? LDS_T : LDS_B64;
    const bool skip_mma = (ti == tj) && (half * T64_N + wn * 32 >= wm * 32 + 32);
    for (int kb = 0; kb < nkb && warp < GEMM_THREADS / 32; kb++) {
        const int s = kb % BULK_STAGES;
        if (BULK_PRODUCER_THREADS == 0 && warp == 0) {      // refill the slot consumed two k-blocks ago with k-block kb + BULK_STAGES - 2
            const int nx = kb + BULK_STAGES - 2;
            if (nx < nkb) {
                if (nx >= BULK_STAGES) bk_mbar_wait(&empty[nx % BULK_STAGES], ((nx / BULK_STAGES) - 1) & 1);
                fill(nx);
            }
        }
        bk_mbar_wait(&full[s], (kb / BULK_STAGES) & 1);
        const double* as = smem + (size_t)s * BULK_STAGE_DOUBLES;
        const double* bs = diag ? (as + half * T64_N) : (as + KB * LDS_T);
        const double* ws = as + KB * (LDS_T + LDS_B64);
        if (!skip_mma) {
            double wcur[KB / 4];
            if (HAS_W) {
#pragma unroll
                for (int kk = 0; kk < KB / 4; kk++) wcur[kk] = ws[kk * 4 + tq];
            }
#pragma unroll
            for (int kk = 0; kk < KB / 4; kk++) {
                double af[4], bf[4];
                const int krow = kk * 4 + tq;
#pragma unroll
                for (int i = 0; i < 4; i++) af[i] = as[krow * LDS_T + wm * 32 + i * 8 + gq];
#pragma unroll
                for (int j = 0; j < 4; j++) { bf[j] = bs[krow * ldsb + wn * 32 + j * 8 + gq]; if (HAS_W) bf[j] *= wcur[kk]; }
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                    for (int j = 0; j < 4; j++) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
            }
        }
        __syncwarp();
        if (lane == 0) bk_arrive(&empty[s]);
    }
    const int rows_valid = g.rows_valid;
    // ---- epilogue through shared memory (Cs[col * TS_LD + row], 64 columns): coalesced 16-byte global accesses
    __syncthreads();
    if (warp < GEMM_THREADS / 32) {
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int e = 0; e < 2; e++) smem[(wn * 32 + j * 8 + tq * 2 + e) * TS_LD + wm * 32 + i * 8 + gq] = acc[i][j][e];
    }
    __syncthreads();
    if (warp >= GEMM_THREADS / 32) return;          // the producer warp has no part in the epilogue (no barrier follows)
    double* __restrict__ C = g.C + (size_t)b * g.strideC;
    const double dinv = (EPI == EPI_ASSEMBLE) ? 1.0 / g.delta[b] : 0.0;
    const double* __restrict__ Pf = (EPI == EPI_ASSEMBLE) ? g.Pf + (size_t)b * g.strideP : nullptr;
    const double* __restrict__ AtA = (EPI == EPI_ASSEMBLE && g.AtA) ? g.AtA + (size_t)b * g.strideAtA : nullptr;
    const double* __restrict__ xr = (EPI == EPI_ASSEMBLE) ? g.xreg + (size_t)b * g.stridex : nullptr;
    const int ncol = g.ncol > 0 ? g.ncol : g.n;
    const int r = rowA0 + (threadIdx.x & 63) * 2;
    constexpr int U = 4;
#pragma unroll 1
    for (int it0 = 0; it0 < T64_N / 4; it0 += U) {
        double2 base[U], extra[U];
        bool ok[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int cl = (it0 + u) * 4 + (threadIdx.x >> 6);
            const int c = rowB0 + cl;
            ok[u] = (r + 1 < g.rows_valid) && (c < ncol) && (r < g.n) && (r + 1 >= c);
            base[u] = make_double2(0.0, 0.0); extra[u] = make_double2(0.0, 0.0);
            if (ok[u]) {
                const size_t idx = (size_t)c * g.ldc + r;
                if (EPI == EPI_SUB) base[u] = *reinterpret_cast<const double2*>(C + idx);
                if (EPI == EPI_ASSEMBLE) { base[u] = *reinterpret_cast<const double2*>(Pf + idx); if (AtA) extra[u] = *reinterpret_cast<const double2*>(AtA + idx); }
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (!ok[u]) continue;
            const int cl = (it0 + u) * 4 + (threadIdx.x >> 6);
            const int c = rowB0 + cl;
            const double2 a = *reinterpret_cast<const double2*>(smem + cl * TS_LD + (threadIdx.x & 63) * 2);
            double v0 = a.x, v1 = a.y;
            if (EPI == EPI_SUB) { v0 = base[u].x - v0; v1 = base[u].y - v1; }
            if (EPI == EPI_ASSEMBLE) {
                double b0 = base[u].x, b1 = base[u].y;
                if (r == c) b0 += xr[r];
                if (r + 1 == c) b1 += xr[r + 1];
                if (AtA) { b0 += dinv * extra[u].x; b1 += dinv * extra[u].y; }
                v0 = b0 + v0; v1 = b1 + v1;
            }
            const size_t idx = (size_t)c * g.ldc + r;
            if (r >= c && r + 1 < g.n) *reinterpret_cast<double2*>(C + idx) = make_double2(v0, v1);
            else { if (r >= c && r < g.n) C[idx] = v0; if (r + 1 >= c && r + 1 < g.n) C[idx + 1] = v1; }
        }
    }
}

#include "dense_chol.cuh"
#include "dense_ozaki.cuh"

// ---------------------------------------------------------------------------------------------------
// trsv: x <- L^{-T} L^{-1} x for one instance per CTA (x in shared memory).
// Forward: right-looking blocks of 32 (coalesced column reads); backward: left-looking (column dots).
// ---------------------------------------------------------------------------------------------------
constexpr int TRSV_THREADS = 512;

// x <- L^{-T} L^{-1} x, one instance per CTA, x in shared memory.  The 32 x 32 diagonal blocks are applied through
// their inverses (by-product of chol_diag_kernel, Linv[blk][k * LB_LD + n] = inv(L_blk)(n, k)), so the serial part of
// each block step is one 32 x 32 mat-vec by warp 0 instead of 32 dependent divisions.
__global__ void __launch_bounds__(TRSV_THREADS, 2)
trsv_kernel(const double* Lmat, long long strideL, int ld, int n, const double* Linv, long long strideLinv,
            double* X, long long strideX, const int* active) {
    extern __shared__ __align__(16) double xs[];   // n doubles + 32 scratch
    const int b = blockIdx.x;
    if (active && !active[b]) return;
    const double* L = Lmat + (size_t)b * strideL;
    const double* Li = Linv + (size_t)b * strideLinv;
    double* x = X + (size_t)b * strideX;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = TRSV_THREADS / 32;
    for (int i = tid; i < n; i += TRSV_THREADS) xs[i] = x[i];
    __syncthreads();
    // ---- forward: L y = x (right-looking, coalesced column reads)
    for (int j0 = 0; j0 < n; j0 += 32) {
        const int nb = min(32, n - j0);
        if (warp == 0) {
            const double* ib = Li + (size_t)(j0 / 32) * LB_SZ;
            const double v = lane < nb ? xs[j0 + lane] : 0.0;
            double y = 0.0;
#pragma unroll
            for (int k8 = 0; k8 < 32; k8 += 8) {
                double col[8];
#pragma unroll
                for (int k = 0; k < 8; k++) col[k] = ib[(k8 + k) * LB_LD + lane];      // inv(lane, k), coalesced
#pragma unroll
                for (int k = 0; k < 8; k++) y += col[k] * __shfl_sync(0xffffffffu, v, k8 + k);
            }
            if (lane < nb) xs[j0 + lane] = y;
        }
        __syncthreads();
        for (int i = j0 + nb + tid; i < n; i += TRSV_THREADS) {
            double s = 0.0;
#pragma unroll 8
            for (int c = 0; c < nb; c++) s += L[(size_t)(j0 + c) * ld + i] * xs[j0 + c];
            xs[i] -= s;
        }
        __syncthreads();
    }
    // ---- backward: L^T x = y (left-looking: column dots over the rows below the block, then inv^T mat-vec)
    double* red = xs + n;   // 32 doubles
    for (int j1 = n; j1 > 0; j1 -= 32) {
        const int j0 = (j1 - 1) / 32 * 32;       // blocks are aligned at multiples of 32
        const int nb = j1 - j0;
        for (int c = warp; c < nb; c += nwarps) {
            const double* col = L + (size_t)(j0 + c) * ld;
            double s = 0.0;
            for (int i = j1 + lane; i < n; i += 32) s += col[i] * xs[i];
            s = warp_sum(s);
            if (lane == 0) red[c] = s;
        }
        __syncthreads();
        if (warp == 0) {
            const double* ib = Li + (size_t)(j0 / 32) * LB_SZ;
            const double v = lane < nb ? xs[j0 + lane] - red[lane] : 0.0;
            double y = 0.0;
#pragma unroll
            for (int k8 = 0; k8 < 32; k8 += 8) {
                double row[8];
#pragma unroll
                for (int k = 0; k < 8; k++) row[k] = ib[lane * LB_LD + k8 + k];      // inv(k, lane)
#pragma unroll
                for (int k = 0; k < 8; k++) y += row[k] * __shfl_sync(0xffffffffu, v, k8 + k);   // sum_k inv(k, lane) v_k
            }
            if (lane < nb) xs[j0 + lane] = y;
        }
        __syncthreads();
        j1 = j0 + 32;   // loop decrement brings it to j0
    }
    for (int i = tid; i < n; i += TRSV_THREADS) x[i] = xs[i];
}

// ---------------------------------------------------------------------------------------------------
// gemv kernels on column-major M (rows x cols, leading dim ld)
//   gemv_n: z[i] (op)= alpha * sum_c M[i,c] * (s ? s[c] : 1) * x[c]      (thread per row, coalesced)
//   gemv_t: z[c]  =  post( alpha * sum_i M[i,c] * x[i] )                 (warp per column, coalesced)
// ---------------------------------------------------------------------------------------------------
struct GemvArgs {
    const double* M; long long strideM; int ld; int rows; int cols;
    const double* x; long long stridex;
    const double* s; long long strides;       // optional per-column (gemv_n) / per-output (gemv_t) scale
    const double* alpha_v;                    // optional per-instance scalar multiplier array (e.g. 1/delta), combined with alpha
    int alpha_v_inverse;                      // use 1/alpha_v[b]
    double alpha;
    double* z; long long stridez;
    int accumulate;                           // gemv_n: z += ... instead of z = ...
    const double* sub; long long stridesub;   // gemv_t: z = s[c] * (alpha*dot - sub[c]*alpha2)
    double alpha2;
    const int* active;
};

// 64 rows per CTA, the columns dealt to 4 thread groups (c = g mod 4) whose partial sums are added in a fixed order: four times
// the loads in flight of the one-thread-per-row version (272 us -> the HBM rate), still deterministic.  Grid: (ceil(rows / 64), batch).
constexpr int GEMV_N_ROWS = 64;
__global__ void __launch_bounds__(256) gemv_n_kernel(GemvArgs a) {
    __shared__ double part[3][GEMV_N_ROWS];
    const int b = blockIdx.y;
    if (a.active && !a.active[b]) return;
    const int r = threadIdx.x & (GEMV_N_ROWS - 1), g = threadIdx.x >> 6;
    const int i = blockIdx.x * GEMV_N_ROWS + r;
    const double* M = a.M + (size_t)b * a.strideM;
    const double* x = a.x + (size_t)b * a.stridex;
    const double* s = a.s ? a.s + (size_t)b * a.strides : nullptr;
    double acc = 0.0;
    if (i < a.rows) {
        if (s) {
#pragma unroll 8
            for (int c = g; c < a.cols; c += 4) acc += M[(size_t)c * a.ld + i] * (x[c] * s[c]);
        } else {
#pragma unroll 8
            for (int c = g; c < a.cols; c += 4) acc += M[(size_t)c * a.ld + i] * x[c];
        }
    }
    if (g > 0) part[g - 1][r] = acc;
    __syncthreads();
    if (g == 0 && i < a.rows) {
        acc = ((acc + part[0][r]) + part[1][r]) + part[2][r];
        double al = a.alpha;
        if (a.alpha_v) al *= a.alpha_v_inverse ? 1.0 / a.alpha_v[b] : a.alpha_v[b];
        double* z = a.z + (size_t)b * a.stridez;
        if (a.accumulate) z[i] += al * acc; else z[i] = al * acc;
    }
}

__global__ void __launch_bounds__(256) gemv_t_kernel(GemvArgs a) {
    const int b = blockIdx.y;
    if (a.active && !a.active[b]) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = blockIdx.x * (blockDim.x >> 5) + warp;
    if (c >= a.cols) return;
    const double* col = a.M + (size_t)b * a.strideM + (size_t)c * a.ld;
    const double* x = a.x + (size_t)b * a.stridex;
    double acc = 0.0;
    for (int i = lane; i < a.rows; i += 32) acc += col[i] * x[i];
    acc = warp_sum(acc);
    if (lane == 0) {
        double al = a.alpha;
        if (a.alpha_v) al *= a.alpha_v_inverse ? 1.0 / a.alpha_v[b] : a.alpha_v[b];
        double v = al * acc;
        if (a.sub) {
            double al2 = a.alpha2;
            if (a.alpha_v) al2 *= a.alpha_v_inverse ? 1.0 / a.alpha_v[b] : a.alpha_v[b];
            v -= al2 * a.sub[(size_t)b * a.stridesub + c];
        }
        if (a.s) v *= a.s[(size_t)b * a.strides + c];
        a.z[(size_t)b * a.stridez + c] = v;
    }
}

#endif  // __CUDACC__
}  // namespace b200
