// piqp_b200/csrc/dense_ozaki.cuh -- K = P + diag(x_reg) + AtA / delta + G^T Z^-1 G on the 5th-generation tensor cores.
//
// Replaces dense::KKT::update_kkt (include/piqp/dense/kkt.hpp:140-160).  tcgen05.mma has no FP64 kind, so the FP64 product
// B B^T (B = G^T Z^-1/2, n x m) is computed EXACTLY in integers (Ozaki splitting) and rounded once:
//
//   oz_rowscale_kernel : e_i = exponent with 2^e_i > max_k |B_ik|   (one power-of-two scale per row of B)
//   oz_split_kernel    : X_ik = rint(B_ik 2^(62 - e_i))  (|X| <= 2^62, int64)  ->  8 balanced base-256 digits
//                        X = sum_a d_a 256^(8-a), d_1 in [-65, 65], d_2..8 in [-128, 127]; digit planes Dg[a][i][k] (s8, k contiguous)
//   oz_gemm_kernel     : per 128 x 64 output tile, sum_k X_ik X_jk = sum_{a,b} 256^(16-a-b) sum_k d_a,ik d_b,jk; the 36 digit pairs with
//                        a + b <= 9 are issued as tcgen05.mma.kind::i8 (s8 x s8 -> s32, exact) into 8 TMEM accumulators, one per weight
//                        g = a + b (8 x 64 columns = the whole TMEM); the dropped pairs are below 2^-58 of the row-scale product.
//                        Operands: one 3-D TMA box {64 B of k, rows, 8 digit planes} per operand and k-step, SWIZZLE_64B, 2-stage
//                        mbarrier ring; warp 0 = TMA producer, warp 1 = MMA issuer (one thread each), then all 4 warps run the
//                        epilogue: tcgen05.ld, Horner recombination of the 8 accumulators in FP64, scale by 2^(e_i + e_j - 68),
//                        add P / diag / AtA, coalesced store of the lower triangle of K.
//
// Measured on this box (tools/umma_i8_probe.cu): the s8 MMA issues at 4.5 POP/s; at this operand intensity the kernel is bound by
// the L2 -> shared-memory operand stream (~6-7 TB/s), which is what the roofline in bench.py reports it against.
// (textually included from dense_kernels.cuh inside namespace b200 and #ifdef __CUDACC__)

constexpr int OZ_S = 8;              // digit planes
constexpr int OZ_F = 62;             // X = rint(B * 2^(OZ_F - e))
constexpr int OZ_TM = 128, OZ_TN = 64;
// two pipeline shapes with the same 192 KB of operand stages: KB = 64 B k-steps x 2 stages (SWIZZLE_64B) or 32 B x 4 stages (SWIZZLE_32B)
constexpr size_t OZ_GEMM_SMEM = (size_t)OZ_S * (OZ_TM + OZ_TN) * 128 + 1024;
constexpr int OZ_EXP_NONFINITE = 0x7fffffff;

struct OzArgs {
    // B_ik = G[(i) + k * ldg] * sqrt(w[k])
    const double* G; long long strideG; int ldg;
    const double* w; long long stridew;
    int n, m, mp;                       // mp = m rounded up to 16 (row pitch of a digit plane in bytes)
    int8_t* Dg;                         // [batch][OZ_S][n][mp]
    int* ex;                            // [batch][n]
    double* sc;                         // [batch][n] 2^(e_i - 34) (NaN when the row holds a non-finite entry): K_ij = V sc_i sc_j
    double* sw;                         // [batch][m] sqrt(w)
    // epilogue
    double* C; long long strideC; int ldc;
    const double* Pf; long long strideP;
    const double* AtA; long long strideAtA;
    const double* xreg; long long stridex;
    const double* delta;
    const int* active;
    const int* tile_ij;                 // [ntiles][2]
    int ntiles;
};

__global__ void oz_sqrt_kernel(const double* w, double* sw, size_t total) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total) sw[i] = sqrt(w[i]);
}
__global__ void __launch_bounds__(128) oz_rowscale_kernel(OzArgs a) {
    const int b = blockIdx.y;
    if (a.active && !a.active[b]) return;
    const int i = blockIdx.x * 128 + threadIdx.x;
    if (i >= a.n) return;
    const double* G = a.G + (size_t)b * a.strideG + i;
    const double* sw = a.sw + (size_t)b * a.stridew;
    double mx0 = 0.0, mx1 = 0.0;
    bool bad = false;
    int k = 0;
    for (; k + 1 < a.m; k += 2) {
        const double v0 = fabs(G[(size_t)k * a.ldg] * sw[k]), v1 = fabs(G[(size_t)(k + 1) * a.ldg] * sw[k + 1]);
        bad |= !(v0 <= 1.7e308) | !(v1 <= 1.7e308);
        mx0 = fmax(mx0, v0); mx1 = fmax(mx1, v1);
    }
    if (k < a.m) { const double v0 = fabs(G[(size_t)k * a.ldg] * sw[k]); bad |= !(v0 <= 1.7e308); mx0 = fmax(mx0, v0); }
    const double mx = fmax(mx0, mx1);
    int e = 0;
    if (mx > 0.0) frexp(mx, &e);                 // mx = f 2^e, f in [0.5, 1)  ->  |B| / 2^e < 1
    a.ex[(size_t)b * a.n + i] = bad ? OZ_EXP_NONFINITE : e;
    a.sc[(size_t)b * a.n + i] = bad ? __longlong_as_double(0x7ff8000000000000ll) : scalbn(1.0, e - (2 * OZ_F - 8 * (OZ_S - 1)) / 2);
}
// block = 256 threads: 32 rows x 128 k.  thread (r = tid % 32, q = tid / 32) converts k = k0 + 16 q .. + 15 of row i0 + r; the digit
// bytes are transposed through shared memory so that every plane row is written as one 128-byte segment.
__global__ void __launch_bounds__(256) oz_split_kernel(OzArgs a) {
    __shared__ __align__(16) unsigned char tile[OZ_S][32][128 + 16];
    const int b = blockIdx.z;
    if (a.active && !a.active[b]) return;
    const int i0 = blockIdx.x * 32, k0 = blockIdx.y * 128;
    const int r = threadIdx.x & 31, q = threadIdx.x >> 5;
    const int i = i0 + r;
    const double* G = a.G + (size_t)b * a.strideG;
    const double* sw = a.sw + (size_t)b * a.stridew;
    int e = 0;
    if (i < a.n) e = a.ex[(size_t)b * a.n + i];
    const bool dead = (i >= a.n) || e == OZ_EXP_NONFINITE;
#pragma unroll 4
    for (int kk = 0; kk < 16; kk++) {
        const int k = k0 + q * 16 + kk;
        long long X = 0;
        if (!dead && k < a.m) X = __double2ll_rn(scalbn(G[(size_t)i + (size_t)k * a.ldg] * sw[k], OZ_F - e));
#pragma unroll
        for (int d = OZ_S - 1; d >= 1; d--) {
            const int dig = (int)((X + 128) & 255) - 128;
            tile[d][r][q * 16 + kk] = (unsigned char)(signed char)dig;
            X = (X - dig) >> 8;
        }
        tile[0][r][q * 16 + kk] = (unsigned char)(signed char)(int)X;
    }
    __syncthreads();
    // write: 8 planes x 32 rows x 8 chunks of 16 B = 2048 chunks
    for (int c = threadIdx.x; c < OZ_S * 32 * 8; c += 256) {
        const int ch = c & 7, rr = (c >> 3) & 31, d = c >> 8;
        const int kb = k0 + ch * 16;
        if (i0 + rr < a.n && kb < a.mp) {
            const uint4 v = *reinterpret_cast<const uint4*>(&tile[d][rr][ch * 16]);
            *reinterpret_cast<uint4*>(a.Dg + (((size_t)b * OZ_S + d) * a.n + (i0 + rr)) * a.mp + kb) = v;
        }
    }
}

// ---- PTX wrappers (tcgen05 / TMA / mbarrier)
__device__ __forceinline__ uint32_t oz_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void oz_mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(oz_smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void oz_mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(oz_smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void oz_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(oz_smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void oz_tma_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(oz_smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(oz_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void oz_umma_s8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void oz_commit(uint64_t* bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(oz_smem_u32(bar)) : "memory"); }
// K-major operand descriptor (cute::UMMA::SmemDescriptor): rows of KB bytes, 8-row groups 8*KB bytes apart, SWIZZLE_64B (4) / SWIZZLE_32B (6)
template <int KB>
__device__ __forceinline__ uint64_t oz_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)((8 * KB) >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)(KB == 64 ? 4 : 6) << 61);
}
__device__ __forceinline__ void oz_tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
}

template <int OZ_KB, int NST>
__global__ void __launch_bounds__(128, 1) oz_gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, OzArgs a) {
    constexpr int OZ_STAGE_A = OZ_S * OZ_TM * OZ_KB, OZ_STAGE_B = OZ_S * OZ_TN * OZ_KB, OZ_STAGE = OZ_STAGE_A + OZ_STAGE_B;
    static_assert((size_t)NST * OZ_STAGE + 1024 <= OZ_GEMM_SMEM, "stages do not fit");
    extern __shared__ __align__(1024) unsigned char oz_sm[];
    __shared__ uint64_t full[NST], empty[NST], done;
    __shared__ uint32_t tmem_base;
    const int b = blockIdx.x / a.ntiles, t = blockIdx.x - b * a.ntiles;
    if (a.active && !a.active[b]) return;
    const int ti = a.tile_ij[2 * t], tj = a.tile_ij[2 * t + 1];
    const int row0 = ti * OZ_TM, col0 = tj * OZ_TN;
    const int tid = threadIdx.x, warp = tid >> 5;
    unsigned char* sm = reinterpret_cast<unsigned char*>(((uintptr_t)oz_sm + 1023) & ~(uintptr_t)1023);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(oz_smem_u32(&tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        for (int s = 0; s < NST; s++) { oz_mbar_init(&full[s], 1); oz_mbar_init(&empty[s], 1); }
        oz_mbar_init(&done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;
    const int nsteps = (a.m + OZ_KB - 1) / OZ_KB;
    if (warp == 0 && (tid & 31) == 0) {                       // ---- TMA producer
        for (int it = 0; it < nsteps; it++) {
            const int s = it % NST;
            if (it >= NST) oz_mbar_wait(&empty[s], ((it / NST) - 1) & 1);
            oz_expect_tx(&full[s], OZ_STAGE);
            oz_tma_3d(sm + (size_t)s * OZ_STAGE, &mapA, it * OZ_KB, row0, b * OZ_S, &full[s]);
            oz_tma_3d(sm + (size_t)s * OZ_STAGE + OZ_STAGE_A, &mapB, it * OZ_KB, col0, b * OZ_S, &full[s]);
        }
    } else if (warp == 1 && (tid & 31) == 0) {                // ---- MMA issuer
        // instruction descriptor (cute::UMMA::InstrDescriptor): D = s32, A = B = s8, K-major, N = 64, M = 128
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(OZ_TN >> 3) << 17) | ((uint32_t)(OZ_TM >> 4) << 24);
        for (int it = 0; it < nsteps; it++) {
            const int s = it % NST;
            oz_mbar_wait(&full[s], (it / NST) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a0 = oz_smem_u32(sm + (size_t)s * OZ_STAGE), b0 = a0 + OZ_STAGE_A;
            // descriptors differ only in the start-address field (bits [0,14) of the low word, units of 16 B): add constants
            const uint64_t da0 = oz_desc<OZ_KB>(a0), db0 = oz_desc<OZ_KB>(b0);
#pragma unroll
            for (int g = 0; g < OZ_S; g++)                    // weight group g = (a + b) - 2, digit indices 0-based below
#pragma unroll
                for (int da = 0; da <= g; da++) {
                    const int db = g - da;
#pragma unroll
                    for (int kk = 0; kk < OZ_KB / 32; kk++)
                        oz_umma_s8(tmem + (uint32_t)(g * OZ_TN), da0 + (uint64_t)((da * (OZ_TM * OZ_KB) + kk * 32) >> 4), db0 + (uint64_t)((db * (OZ_TN * OZ_KB) + kk * 32) >> 4), idesc,
                                   (it > 0 || da > 0 || kk > 0) ? 1u : 0u);
                }
            oz_commit(&empty[s]);
        }
        oz_commit(&done);
    }
    // ---- epilogue (all 4 warps; thread = row of the tile)
    const int gi = row0 + tid;
    double* C = a.C + (size_t)b * a.strideC;
    const double* Pf = a.Pf + (size_t)b * a.strideP;
    const double* AtA = a.AtA ? a.AtA + (size_t)b * a.strideAtA : nullptr;
    const double dinv = a.AtA ? 1.0 / a.delta[b] : 0.0;
    const double* sc = a.sc + (size_t)b * a.n;
    const double si = gi < a.n ? sc[gi] : 0.0;
    const double xr = gi < a.n ? a.xreg[(size_t)b * a.stridex + gi] : 0.0;
    auto load_base = [&](int c0, double (&base)[8], double (&sj)[8]) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int gj = col0 + c0 + j;
            base[j] = 0.0; sj[j] = 0.0;
            if (gi < a.n && gj <= gi) { const size_t idx = (size_t)gj * a.ldc + gi; base[j] = Pf[idx]; if (AtA) base[j] += dinv * AtA[idx]; sj[j] = sc[gj]; }
        }
    };
    double base[8], sj[8];
    load_base(0, base, sj);                       // in flight while the last MMAs run
    oz_mbar_wait(&done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
    for (int c0 = 0; c0 < OZ_TN; c0 += 8) {
        uint32_t r[OZ_S][8];
#pragma unroll
        for (int g = 0; g < OZ_S; g++) oz_tmem_ld8(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(g * OZ_TN + c0), r[g]);
        double nbase[8], nsj[8];
        if (c0 + 8 < OZ_TN) load_base(c0 + 8, nbase, nsj);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int gj = col0 + c0 + j;
            if (gi < a.n && gj <= gi) {
                double V = (double)(int)r[0][j];
#pragma unroll
                for (int g = 1; g < OZ_S; g++) V = fma(V, 256.0, (double)(int)r[g][j]);
                double bse = base[j];
                if (gj == gi) bse += xr;
                C[(size_t)gj * a.ldc + gi] = bse + (V * si) * sj[j];
            }
        }
#pragma unroll
        for (int j = 0; j < 8; j++) { base[j] = nbase[j]; sj[j] = nsj[j]; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}
