// piqp_b200/csrc/dense_chol.cuh -- batched blocked Cholesky (replaces Eigen::LLT::compute, dense/kkt.hpp:82).
//
// Left-looking over 128-column block columns; two kernels per block column jb:
//   chol_diag_kernel  : one CTA per instance.  Diagonal tile -> shared memory, blocked (32) Cholesky in place:
//                       DMMA for the inter-block updates and the below-diagonal solves, one warp for the 32 x 32
//                       pivot blocks, and the inverses of the four 32 x 32 diagonal blocks of L11 as a by-product.
//   chol_panel_kernel : one CTA per (instance, row tile below).  DMMA mainloop for the left-looking update
//                       T = A - sum_k L(ti,k) L(jb,k)^T, then the triangular solve X L11^T = T as a blocked forward
//                       substitution on 16-row strips (one warp each, all DMMA: off-diagonal blocks of L11 and the
//                       inverted diagonal blocks), then the rank-128 update of this row tile's OWN diagonal tile
//                       (right-looking look-ahead), so the diag kernel has no mainloop.
// All fp64; pivots <= 0 (or NaN) report failure like Eigen's LLT (info() != Success).
// (textually included from dense_kernels.cuh inside namespace b200 and #ifdef __CUDACC__)

constexpr int CB = 32;                 // inner block
constexpr int LB_LD = CB + 4;          // leading dim of a 32 x 32 block in smem: Bs[k * LB_LD + n]
constexpr int LB_SZ = CB * LB_LD;      // doubles per block
constexpr size_t CHOL_PANEL_SMEM = TILE_SMEM + 10 * LB_SZ * sizeof(double);
constexpr int CHOL_THREADS = 256;

// acc(16 x 32) += sgn * A(16 x K) * B(K x 32); As[k * lda + row], Bs[k * ldb + n]; K multiple of 4
__device__ __forceinline__ void warp_mma_16x32(double (&acc)[2][4][2], const double* As, int lda, const double* Bs, int ldb, int K, bool negate) {
    const int lane = threadIdx.x & 31, gq = lane >> 2, tq = lane & 3;
#pragma unroll 4
    for (int k0 = 0; k0 < K; k0 += 4) {
        double a[2], bb[4];
#pragma unroll
        for (int i = 0; i < 2; i++) { a[i] = As[(k0 + tq) * lda + i * 8 + gq]; if (negate) a[i] = -a[i]; }
#pragma unroll
        for (int j = 0; j < 4; j++) bb[j] = Bs[(k0 + tq) * ldb + j * 8 + gq];
#pragma unroll
        for (int i = 0; i < 2; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], bb[j]);
    }
}
// 16 x 32 block in C-fragment layout <-> Ts (element (row, col) at Ts[col * TS_LD + row])
__device__ __forceinline__ void cfrag_load(double (&acc)[2][4][2], const double* Ts) {
    const int lane = threadIdx.x & 31, gq = lane >> 2, tq = lane & 3;
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int e = 0; e < 2; e++) acc[i][j][e] = Ts[(j * 8 + tq * 2 + e) * TS_LD + i * 8 + gq];
}
__device__ __forceinline__ void cfrag_store(const double (&acc)[2][4][2], double* Ts) {
    const int lane = threadIdx.x & 31, gq = lane >> 2, tq = lane & 3;
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int e = 0; e < 2; e++) Ts[(j * 8 + tq * 2 + e) * TS_LD + i * 8 + gq] = acc[i][j][e];
}
__device__ __forceinline__ void cfrag_zero(double (&acc)[2][4][2]) {
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
}

// One warp: Cholesky of the 32 x 32 block T (T[c * TS_LD + r], lower) in place and inv <- L^{-1} (inv[k * LB_LD + n] = Linv(n,k)).
// Returns 0 or (failing column + 1).
__device__ __noinline__ int warp_potf2_32(double* T, double* inv) {
    const int lane = threadIdx.x & 31;
    double a[CB];
#pragma unroll
    for (int c = 0; c < CB; c++) a[c] = (c <= lane) ? T[c * TS_LD + lane] : 0.0;
    int failed = 0;
    double my_rinv = 1.0;           // lane k keeps 1 / l_kk
#pragma unroll
    for (int k = 0; k < CB; k++) {
        double dk = __shfl_sync(0xffffffffu, a[k], k);
        if (!(dk > 0.0)) { if (!failed) failed = k + 1; dk = 1.0; }
        const double rinv = rsqrt(dk);
        const double lkk = dk * rinv;
        const double lrk = a[k] * rinv;
        a[k] = (lane == k) ? lkk : lrk;
        if (lane == k) my_rinv = rinv;
#pragma unroll
        for (int j = k + 1; j < CB; j++) {
            const double ljk = __shfl_sync(0xffffffffu, lrk, j);
            if (lane >= j) a[j] -= lrk * ljk;
        }
    }
#pragma unroll
    for (int c = 0; c < CB; c++) if (c <= lane) T[c * TS_LD + lane] = a[c];
    __syncwarp();
    // inverse, column `lane` of L^{-1}: x_i = ((i == lane) - sum_{k<i} l_ik x_k) / l_ii   (x_k = 0 for k < lane)
    double x[CB];
#pragma unroll
    for (int i = 0; i < CB; i++) {
        double s = (i == lane) ? 1.0 : 0.0;
#pragma unroll
        for (int k = 0; k < i; k++) s -= T[k * TS_LD + i] * x[k];
        x[i] = s * __shfl_sync(0xffffffffu, my_rinv, i);
    }
#pragma unroll
    for (int i = 0; i < CB; i++) inv[lane * LB_LD + i] = x[i];
    return failed;
}

// ---------------------------------------------------------------------------------------------------
// The 128 x 128 diagonal tile is held as its 10 lower 32 x 32 BLOCKS (block (bi, bk), bk <= bi, at index bi (bi + 1) / 2 + bk, element
// (r, c) at [c * LB_LD + r]): 92 KB + one inverse block = 101 KB and <= 128 registers, so TWO diag CTAs share an SM and the 256
// instances of BASELINE config 2 are one wave on 148 SMs instead of two (the kernel is a latency chain with one busy warp).
constexpr size_t CHOL_DIAG_SMEM = (size_t)11 * LB_SZ * sizeof(double);
__device__ __forceinline__ double* cd_blk(double* Tb, int bi, int bk) { return Tb + (size_t)(bi * (bi + 1) / 2 + bk) * LB_SZ; }
__device__ __forceinline__ void cfrag_load_ld(double (&acc)[2][4][2], const double* Ts, int ldt);
__device__ __forceinline__ void cfrag_store_ld(const double (&acc)[2][4][2], double* Ts, int ldt);
// one warp: Cholesky of a 32 x 32 block stored with leading dimension LB_LD (wrapper around warp_potf2_32's TS_LD layout is avoided
// by templating the leading dimension)
template <int LDT>
__device__ __noinline__ int warp_potf2_32_ld(double* T, double* inv) {
    const int lane = threadIdx.x & 31;
    double a[CB];
#pragma unroll
    for (int c = 0; c < CB; c++) a[c] = (c <= lane) ? T[c * LDT + lane] : 0.0;
    int failed = 0;
    double my_rinv = 1.0;
#pragma unroll
    for (int k = 0; k < CB; k++) {
        double dk = __shfl_sync(0xffffffffu, a[k], k);
        if (!(dk > 0.0)) { if (!failed) failed = k + 1; dk = 1.0; }
        const double rinv = rsqrt(dk);
        const double lkk = dk * rinv;
        const double lrk = a[k] * rinv;
        a[k] = (lane == k) ? lkk : lrk;
        if (lane == k) my_rinv = rinv;
#pragma unroll
        for (int j = k + 1; j < CB; j++) {
            const double ljk = __shfl_sync(0xffffffffu, lrk, j);
            if (lane >= j) a[j] -= lrk * ljk;
        }
    }
#pragma unroll
    for (int c = 0; c < CB; c++) if (c <= lane) T[c * LDT + lane] = a[c];
    __syncwarp();
    double x[CB];
#pragma unroll
    for (int i = 0; i < CB; i++) {
        double s = (i == lane) ? 1.0 : 0.0;
#pragma unroll
        for (int k = 0; k < i; k++) s -= T[k * LDT + i] * x[k];
        x[i] = s * __shfl_sync(0xffffffffu, my_rinv, i);
    }
#pragma unroll
    for (int i = 0; i < CB; i++) inv[lane * LB_LD + i] = x[i];
    return failed;
}
__global__ void __launch_bounds__(CHOL_THREADS, 2)
chol_diag_kernel(double* Kmat, long long strideK, int ld, int n, int j0, double* Linv, long long strideLinv, int* fail, const int* active) {
    extern __shared__ __align__(16) double smem[];
    double* Tb = smem;                          // 10 lower blocks of the tile
    double* inv = smem + 10 * LB_SZ;            // ONE 32 x 32 inverse block at a time (written to global memory as soon as the rows below have used it)
    __shared__ int s_fail;
    const int b = blockIdx.x;
    if (active && !active[b]) return;
    if (fail[b]) return;
    double* K = Kmat + (size_t)b * strideK;
    const int nb = min(TILE, n - j0);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) s_fail = 0;
    // lower triangle of the tile; identity on the padding so that partial tiles factor trivially there
    for (int blk = 0; blk < 10; blk++) {
        int bi = 0; while ((bi + 1) * (bi + 2) / 2 <= blk) bi++;
        const int bk = blk - bi * (bi + 1) / 2;
        double* B = Tb + (size_t)blk * LB_SZ;
        for (int c = warp; c < CB; c += CHOL_THREADS / 32) {
            const int gr = bi * CB + lane, gc = bk * CB + c;
            double v = (gr == gc) ? 1.0 : 0.0;
            if (gr < nb && gc < nb) v = (gr >= gc) ? K[(size_t)(j0 + gc) * ld + j0 + gr] : 0.0;
            B[c * LB_LD + lane] = v;
        }
    }
    double* ib = Linv + (size_t)b * strideLinv + (size_t)(j0 / CB) * LB_SZ;   // blocks j0/32 .. j0/32 + 3 (allocation is padded to whole tiles)
    __syncthreads();
    const int nblk = (nb + CB - 1) / CB;
    const int bi_w = warp >> 1, ro = (warp & 1) * 16;      // this warp's 16-row strip: block row bi_w, rows ro .. ro + 15 inside the block
    for (int s = 0; s < nblk; s++) {
        // (1) left-looking update of block column s with block columns < s, rows >= 32 s
        if (s > 0 && bi_w >= s) {
            double acc[2][4][2];
            cfrag_load_ld(acc, cd_blk(Tb, bi_w, s) + ro, LB_LD);
            for (int k = 0; k < s; k++)      // A = T(strip, block column k); B(kk, nn) = T(32 s + nn, 32 k + kk) = block (s, k) [kk * LB_LD + nn]
                warp_mma_16x32(acc, cd_blk(Tb, bi_w, k) + ro, LB_LD, cd_blk(Tb, s, k), LB_LD, CB, true);
            __syncwarp();
            cfrag_store_ld(acc, cd_blk(Tb, bi_w, s) + ro, LB_LD);
        }
        for (int i = tid; i < LB_SZ; i += CHOL_THREADS) inv[i] = 0.0;      // the previous block's inverse went to global memory before the last barrier
        __syncthreads();
        // (2) pivot block + its inverse
        if (warp == 0) {
            const int f = warp_potf2_32_ld<LB_LD>(cd_blk(Tb, s, s), inv);
            if (f && (tid == 0)) s_fail = j0 + CB * s + f;
        }
        __syncthreads();
        if (s_fail) break;
        // (3) rows below: X = T * L_ss^{-T}
        if (bi_w >= s + 1) {
            double acc[2][4][2];
            cfrag_zero(acc);
            warp_mma_16x32(acc, cd_blk(Tb, bi_w, s) + ro, LB_LD, inv, LB_LD, CB, false);
            __syncwarp();
            cfrag_store_ld(acc, cd_blk(Tb, bi_w, s) + ro, LB_LD);
        }
        for (int i = tid; i < LB_SZ; i += CHOL_THREADS) ib[(size_t)s * LB_SZ + i] = inv[i];
        __syncthreads();
    }
    if (s_fail) { if (tid == 0) fail[b] = s_fail; return; }
    for (int s = nblk; s < 4; s++) for (int i = tid; i < LB_SZ; i += CHOL_THREADS) ib[(size_t)s * LB_SZ + i] = 0.0;
    for (int blk = 0; blk < 10; blk++) {
        int bi = 0; while ((bi + 1) * (bi + 2) / 2 <= blk) bi++;
        const int bk = blk - bi * (bi + 1) / 2;
        const double* B = Tb + (size_t)blk * LB_SZ;
        for (int c = warp; c < CB; c += CHOL_THREADS / 32) {
            const int gr = bi * CB + lane, gc = bk * CB + c;
            if (gr < nb && gc < nb && gr >= gc) K[(size_t)(j0 + gc) * ld + j0 + gr] = B[c * LB_LD + lane];
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// FUSED = true: the round-1 kernel (left-looking mainloop + solve + look-ahead update of the own diagonal tile in one CTA that holds a
// whole SM).  FUSED = false (default since round 2): solve only -- the left-looking update of the whole block column (diagonal tile
// included) is done beforehand by gemm_nt_t64_kernel<EPI_SUB> (two CTAs per SM, no serial tail behind its mainloop).
template <bool FUSED>
__global__ void __launch_bounds__(CHOL_THREADS, 1)
chol_panel_kernel(double* Kmat, long long strideK, int ld, int n, int jb, int row_tiles, const double* Linv, long long strideLinv, const int* fail, const int* active) {
    extern __shared__ __align__(16) double smem[];
    double* Ts = smem;                          // the tile being solved: Ts[col * TS_LD + row]
    double* Lb = smem + TILE * TS_LD;           // 6 off-diagonal 32 x 32 blocks of L11, then 4 inverse diagonal blocks
    double* inv = Lb + 6 * LB_SZ;
    const int b = blockIdx.x / row_tiles, t = blockIdx.x % row_tiles;
    if (active && !active[b]) return;
    if (fail[b]) return;
    double* K = Kmat + (size_t)b * strideK;
    const int j0 = jb * TILE, ti = jb + 1 + t, rows0 = ti * TILE;
    const int tid = threadIdx.x, warp = tid >> 5;

    // ---- prefetch the blocks of L11 (Lb(s,k)[kk * LB_LD + nn] = L11(32 s + nn, 32 k + kk), k < s) and the inverse
    //      diagonal blocks with cp.async into the smem region behind the mainloop stages; they land during the mainloop
    for (int e = tid; e < 6 * CB * (CB / 2); e += CHOL_THREADS) {
        const int blk = e / (CB * CB / 2), kk = (e / (CB / 2)) % CB, n2 = (e % (CB / 2)) * 2;
        int s = 1, k = blk;
        while (k >= s) { k -= s; s++; }
        cp_async16(Lb + blk * LB_SZ + kk * LB_LD + n2, K + (size_t)(j0 + CB * k + kk) * ld + j0 + CB * s + n2, 16);
    }
    {
        const double* ib = Linv + (size_t)b * strideLinv + (size_t)(j0 / CB) * LB_SZ;
        for (int i = tid; i < 4 * LB_SZ / 2; i += CHOL_THREADS) cp_async16(inv + 2 * i, ib + 2 * i, 16);
    }
    cp_async_commit();

    // ---- left-looking update: acc = sum_{k < j0} L(ti,k) L(jb,k)^T
    double acc[8][4][2];
    if (FUSED) {
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
        if (j0 > 0) {
            gemm_mainloop<false>(acc, smem, K, ld, rows0, K, ld, j0, ld, j0, nullptr, false);
            __syncthreads();
        }
        acc_to_smem(acc, Ts, -1.0);
    } else {
        for (int e = tid; e < TILE * TS_LD; e += CHOL_THREADS) Ts[e] = 0.0;
    }
    cp_async_wait<0>();
    __syncthreads();
    // ---- T = A - acc: all 32 16-byte loads of this thread are in flight before the first use
    {
        const int r = (tid & 63) * 2;
        const bool rok = rows0 + r + 1 < ld + 0;
        double2 a[TILE / 4];
#pragma unroll
        for (int it = 0; it < TILE / 4; it++) {
            const int c = it * 4 + (tid >> 6);
            a[it] = make_double2(0.0, 0.0);
            if (rok) a[it] = *reinterpret_cast<const double2*>(K + (size_t)(j0 + c) * ld + rows0 + r);
        }
#pragma unroll
        for (int it = 0; it < TILE / 4; it++) {
            const int c = it * 4 + (tid >> 6);
            double2* p = reinterpret_cast<double2*>(Ts + c * TS_LD + r);
            double2 v = *p; v.x += a[it].x; v.y += a[it].y; *p = v;
        }
    }
    __syncthreads();
    // ---- X * L11^T = T on this warp's 16-row strip, blocked forward substitution (no block-level barriers)
    {
        const int row0 = warp * 16;
        for (int s = 0; s < 4; s++) {
            double c2[2][4][2];
            if (s > 0) {
                cfrag_load(c2, Ts + (CB * s) * TS_LD + row0);
                for (int k = 0; k < s; k++)
                    warp_mma_16x32(c2, Ts + (CB * k) * TS_LD + row0, TS_LD, Lb + (s * (s - 1) / 2 + k) * LB_SZ, LB_LD, CB, true);
                __syncwarp();
                cfrag_store(c2, Ts + (CB * s) * TS_LD + row0);
                __syncwarp();
            }
            cfrag_zero(c2);
            warp_mma_16x32(c2, Ts + (CB * s) * TS_LD + row0, TS_LD, inv + s * LB_SZ, LB_LD, CB, false);
            __syncwarp();
            cfrag_store(c2, Ts + (CB * s) * TS_LD + row0);
            __syncwarp();
        }
    }
    __syncthreads();
    // ---- store L21 tile
    {
        const int r = (tid & 63) * 2;
#pragma unroll 4
        for (int it = 0; it < TILE / 4; it++) {
            const int c = it * 4 + (tid >> 6);
            const double2 v = *reinterpret_cast<const double2*>(Ts + c * TS_LD + r);
            double* dst = K + (size_t)(j0 + c) * ld + rows0 + r;
            if (rows0 + r + 1 < n) *reinterpret_cast<double2*>(dst) = v;
            else if (rows0 + r < n) dst[0] = v.x;
        }
    }
    // ---- look-ahead: this row tile's own diagonal tile (ti,ti) gets its rank-128 contribution of block column jb now
    //      (D -= X X^T, lower part), so chol_diag_kernel never needs a mainloop and the work is spread over all panel CTAs
    if (FUSED) {
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
        const int lane = tid & 31, wm = warp >> 2, wn = warp & 3, gq = lane >> 2, tq = lane & 3;
#pragma unroll 2
        for (int k0 = 0; k0 < TILE; k0 += 4) {
            double af[8], bf[4];
#pragma unroll
            for (int i = 0; i < 8; i++) af[i] = Ts[(k0 + tq) * TS_LD + wm * 64 + i * 8 + gq];
#pragma unroll
            for (int j = 0; j < 4; j++) bf[j] = Ts[(k0 + tq) * TS_LD + wn * 32 + j * 8 + gq];
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
        __syncthreads();
        acc_to_smem(acc, Ts, 1.0);
        __syncthreads();
        const int r = (tid & 63) * 2;
        const int d0 = rows0;                     // this row tile's diagonal tile starts at row/col rows0
        double2 old[TILE / 4];
#pragma unroll
        for (int it = 0; it < TILE / 4; it++) {
            const int c = it * 4 + (tid >> 6);
            const int gr = d0 + r, gc = d0 + c;
            old[it] = make_double2(0.0, 0.0);
            if (gc < n && gr + 1 >= gc && gr + 1 < ld) old[it] = *reinterpret_cast<const double2*>(K + (size_t)gc * ld + gr);
        }
#pragma unroll
        for (int it = 0; it < TILE / 4; it++) {
            const int c = it * 4 + (tid >> 6);
            const int gr = d0 + r, gc = d0 + c;
            if (gc >= n || gr + 1 < gc || gr >= n) continue;
            const double2 u = *reinterpret_cast<const double2*>(Ts + c * TS_LD + r);
            double* dst = K + (size_t)gc * ld + gr;
            if (gr >= gc && gr + 1 < n) *reinterpret_cast<double2*>(dst) = make_double2(old[it].x - u.x, old[it].y - u.y);
            else { if (gr >= gc && gr < n) dst[0] = old[it].x - u.x; if (gr + 1 >= gc && gr + 1 < n) dst[1] = old[it].y - u.y; }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// Panel solve on 64-row half tiles (default): X L11^T = T for rows [rows0, rows0 + 64) of block column jb.  Only the T tile lives in
// shared memory (68 KB): the off-diagonal 32 x 32 blocks of L11 and the inverted diagonal blocks are read as DMMA B operands straight
// from global memory (they are shared by every CTA of the instance and L2-hot), so three CTAs fit one SM -- or one beside a
// tile-kernel CTA -- instead of one CTA holding a whole SM behind a latency-bound substitution.  4 warps, one 16-row strip each.
constexpr int CS_ROWS = 64;
constexpr int CS_THREADS = 128;
constexpr int CS_LD = CS_ROWS + 4;
constexpr size_t CHOL_SOLVE64_SMEM = (size_t)TILE * CS_LD * sizeof(double);
__device__ __forceinline__ void cfrag_load_ld(double (&acc)[2][4][2], const double* Ts, int ldt) {
    const int lane = threadIdx.x & 31, gq = lane >> 2, tq = lane & 3;
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int e = 0; e < 2; e++) acc[i][j][e] = Ts[(j * 8 + tq * 2 + e) * ldt + i * 8 + gq];
}
__device__ __forceinline__ void cfrag_store_ld(const double (&acc)[2][4][2], double* Ts, int ldt) {
    const int lane = threadIdx.x & 31, gq = lane >> 2, tq = lane & 3;
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int e = 0; e < 2; e++) Ts[(j * 8 + tq * 2 + e) * ldt + i * 8 + gq] = acc[i][j][e];
}
__global__ void __launch_bounds__(CS_THREADS, 3)
chol_solve64_kernel(double* Kmat, long long strideK, int ld, int n, int jb, int half_tiles, const double* Linv, long long strideLinv, const int* fail, const int* active) {
    extern __shared__ __align__(16) double smem[];
    double* Ts = smem;                          // Ts[col * CS_LD + row], 128 columns x 64 rows
    const int b = blockIdx.x / half_tiles, t = blockIdx.x % half_tiles;
    if (active && !active[b]) return;
    if (fail[b]) return;
    double* K = Kmat + (size_t)b * strideK;
    const int j0 = jb * TILE, rows0 = (jb + 1) * TILE + t * CS_ROWS;
    if (rows0 >= n) return;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int r = (tid & 31) * 2, cg = tid >> 5;                 // two rows per thread, 4 column groups
    {
        const bool rok = rows0 + r + 1 < ld;
        double2 a[TILE / 4];
#pragma unroll
        for (int it = 0; it < TILE / 4; it++) {
            const int c = it * 4 + cg;
            a[it] = make_double2(0.0, 0.0);
            if (rok) a[it] = *reinterpret_cast<const double2*>(K + (size_t)(j0 + c) * ld + rows0 + r);
        }
#pragma unroll
        for (int it = 0; it < TILE / 4; it++) { const int c = it * 4 + cg; *reinterpret_cast<double2*>(Ts + c * CS_LD + r) = a[it]; }
    }
    __syncthreads();
    {
        const int row0 = warp * 16;
        const double* L11 = K + (size_t)j0 * ld + j0;                                   // L11(i, k) at L11[k * ld + i]
        const double* inv = Linv + (size_t)b * strideLinv + (size_t)(j0 / CB) * LB_SZ;  // block s: inv[s * LB_SZ + k * LB_LD + nn] = Linv_ss(nn, k)
        for (int s = 0; s < 4; s++) {
            double c2[2][4][2];
            if (s > 0) {
                cfrag_load_ld(c2, Ts + (CB * s) * CS_LD + row0, CS_LD);
                for (int k = 0; k < s; k++)      // B(kk, nn) = L11(32 s + nn, 32 k + kk)
                    warp_mma_16x32(c2, Ts + (CB * k) * CS_LD + row0, CS_LD, L11 + (size_t)(CB * k) * ld + CB * s, ld, CB, true);
                __syncwarp();
                cfrag_store_ld(c2, Ts + (CB * s) * CS_LD + row0, CS_LD);
                __syncwarp();
            }
            cfrag_zero(c2);
            warp_mma_16x32(c2, Ts + (CB * s) * CS_LD + row0, CS_LD, inv + (size_t)s * LB_SZ, LB_LD, CB, false);
            __syncwarp();
            cfrag_store_ld(c2, Ts + (CB * s) * CS_LD + row0, CS_LD);
            __syncwarp();
        }
    }
    __syncthreads();
#pragma unroll 4
    for (int it = 0; it < TILE / 4; it++) {
        const int c = it * 4 + cg;
        const double2 v = *reinterpret_cast<const double2*>(Ts + c * CS_LD + r);
        double* dst = K + (size_t)(j0 + c) * ld + rows0 + r;
        if (rows0 + r + 1 < n) *reinterpret_cast<double2*>(dst) = v;
        else if (rows0 + r < n) dst[0] = v.x;
    }
}
