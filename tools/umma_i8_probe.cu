// tools/umma_i8_probe.cu -- does the 5th-generation tensor core (tcgen05.mma.kind::i8, accumulators in TMEM, operands staged by
// TMA) work on this box, and how fast is it?  Groundwork for the Ozaki-split FP64 assembly GEMM.
//
//   part A  correctness: one CTA, D[128 x N] (s32) = A[128 x K] (s8) * B[N x K]^T (u8 or s8), K-major operands written to
//           shared memory in the SWIZZLE_128B canonical layout by plain stores, result read back with tcgen05.ld and
//           compared with a host loop.  Also checks accumulation over several issues and the mixed s8 x u8 formats the
//           two's-complement slicing needs.
//   part B  issue-rate ceiling: 148 CTAs re-issue MMAs on resident operands (no loads): TOP/s for N = 64 / 128 / 256.
//   part C  TMA-fed stream: operands pulled from L2/HBM by cp.async.bulk.tensor (2 stages, mbarrier pipeline) at the
//           operand intensity of the Ozaki kernel (P slice-pair MMAs per loaded 128-byte k-slab).
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/umma_i8_probe tools/umma_i8_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) { asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory"); }
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
                   "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// shared-memory matrix descriptor, K-major, SWIZZLE_128B: rows of 128 bytes, 8-row groups 1024 bytes apart
// (cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), layout_type [61,64))
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;                    // LBO (unused for swizzled K-major) = 1
    d |= (uint64_t)(1024 >> 4) << 32;          // SBO = 1024 B
    d |= (uint64_t)1 << 46;                    // version 1 (Blackwell)
    d |= (uint64_t)2 << 61;                    // SWIZZLE_128B
    return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): c_format S32 = 2 [4,6), a_format [7,10), b_format [10,13) (0 = u8, 1 = s8),
// a/b major K = 0, N>>3 [17,23), M>>4 [24,29)
__host__ __device__ inline uint32_t make_idesc_i8(int M, int N, int a_signed, int b_signed) {
    return (2u << 4) | ((uint32_t)a_signed << 7) | ((uint32_t)b_signed << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// byte offset of element (row r, byte k) of a [rows x 128 B] slab in the SWIZZLE_128B K-major layout
__host__ __device__ inline int sw128_off(int r, int k) { return (r >> 3) * 1024 + (r & 7) * 128 + ((((k >> 4) ^ (r & 7)) & 7) << 4) + (k & 15); }

// ------------------------------------------------------------------------------------------------ part A
// A: [128][K] s8 row-major (global), B: [N][K] row-major (global), D: [128][N] s32.  K multiple of 128, N multiple of 16 <= 256.
__global__ void __launch_bounds__(128) probe_correct(const int8_t* A, const uint8_t* B, int32_t* D, int N, int K, int b_signed, int splits) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int nslab = K / 128;
    uint8_t* sA = sm;                              // nslab slabs of 128 x 128 B
    uint8_t* sB = sm + (size_t)nslab * 128 * 128;   // nslab slabs of N x 128 B
    for (int e = tid; e < 128 * K; e += 128) { const int r = e / K, k = e % K; sA[(k / 128) * 128 * 128 + sw128_off(r, k % 128)] = (uint8_t)A[e]; }
    for (int e = tid; e < N * K; e += 128) { const int r = e / K, k = e % K; sB[(k / 128) * N * 128 + sw128_off(r, k % 128)] = B[e]; }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy stores -> visible to the tensor core (async proxy)
    if (warp == 0) tmem_alloc(&tmem_base, 256);
    if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base;
    if (tid == 0) {
        const uint32_t idesc = make_idesc_i8(128, N, 1, b_signed);
        // `splits` separate commit groups to check accumulation across issues
        for (int s = 0; s < nslab; s++)
            for (int kk = 0; kk < 4; kk++) {
                const uint64_t ad = make_desc_sw128(smem_u32(sA + (size_t)s * 128 * 128) + kk * 32);
                const uint64_t bd = make_desc_sw128(smem_u32(sB + (size_t)s * N * 128) + kk * 32);
                umma_i8(tmem, ad, bd, idesc, (s | kk) ? 1u : 0u);
            }
        umma_commit(&bar);
    }
    (void)splits;
    mbar_wait(&bar, 0);
    tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, r);
        for (int j = 0; j < 16; j++) D[(size_t)tid * N + c0 + j] = (int32_t)r[j];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

// ------------------------------------------------------------------------------------------------ part A2
// The layout the Ozaki kernel uses: operands as S digit slices [slice][row][K] in global memory, pulled by ONE 3-D TMA box
// {64 B of k, rows, S slices} with SWIZZLE_64B into [slice][row][64 B] shared memory, consumed through SWIZZLE_64B descriptors.
// D_g[128 x 64] = sum over slice pairs (a, b) with a + b = g of A_a B_b^T, g = 0 .. 2S-2 restricted to g < S  -> S accumulators.
__device__ __forceinline__ uint64_t make_desc_sw64(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(512 >> 4) << 32;           // SBO = 8 rows x 64 B
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;                    // SWIZZLE_64B
    return d;
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}
template <int S>
__global__ void __launch_bounds__(128) probe_sliced(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, int32_t* D, int K, int rowA0, int rowB0) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t full[2], empty[2], done;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    constexpr int SA = S * 128 * 64, SB = S * 64 * 64, STAGE = SA + SB;
    if (warp == 0) tmem_alloc(&tmem_base, 512);
    if (tid == 0) { for (int s = 0; s < 2; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); } mbar_init(&done, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base;
    const int nsteps = (K + 63) / 64;
    if (warp == 0 && (tid & 31) == 0) {
        for (int it = 0; it < nsteps; it++) {
            const int s = it & 1;
            if (it >= 2) mbar_wait(&empty[s], ((it >> 1) - 1) & 1);
            mbar_expect_tx(&full[s], STAGE);
            tma_load_3d(sm + (size_t)s * STAGE, &mapA, it * 64, rowA0, 0, &full[s]);
            tma_load_3d(sm + (size_t)s * STAGE + SA, &mapB, it * 64, rowB0, 0, &full[s]);
        }
    } else if (warp == 1 && (tid & 31) == 0) {
        const uint32_t idesc = make_idesc_i8(128, 64, 1, 1);
        for (int it = 0; it < nsteps; it++) {
            const int s = it & 1;
            mbar_wait(&full[s], (it >> 1) & 1);
            tc_fence_after();
            const uint32_t a0 = smem_u32(sm + (size_t)s * STAGE), b0 = a0 + SA;
            for (int g = 0; g < S; g++)
                for (int a = 0; a <= g; a++) {
                    const int b2 = g - a;
#pragma unroll
                    for (int kk = 0; kk < 2; kk++)
                        umma_i8(tmem + (uint32_t)(g * 64), make_desc_sw64(a0 + a * (128 * 64) + kk * 32), make_desc_sw64(b0 + b2 * (64 * 64) + kk * 32), idesc,
                                (it > 0 || a > 0 || kk > 0) ? 1u : 0u);
                }
            umma_commit(&empty[s]);
        }
        umma_commit(&done);
    }
    mbar_wait(&done, 0);
    tc_fence_after();
    for (int g = 0; g < S; g++)
        for (int c0 = 0; c0 < 64; c0 += 16) {
            uint32_t r[16];
            tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + g * 64 + c0, r);
            for (int j = 0; j < 16; j++) D[((size_t)g * 128 + tid) * 64 + c0 + j] = (int32_t)r[j];
        }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------ part B
__global__ void __launch_bounds__(128) probe_issue(int N, int iters, unsigned long long* cycles_out) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int e = tid; e < (128 + 256) * 128 / 4; e += 128) reinterpret_cast<uint32_t*>(sm)[e] = 0x01010101u * (uint32_t)(e & 3);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) tmem_alloc(&tmem_base, 512);
    if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base;
    unsigned long long t0 = 0, t1 = 0;
    if (tid == 0) {
        const uint32_t idesc = make_idesc_i8(128, N, 1, 0);
        const uint32_t a0 = smem_u32(sm), b0 = smem_u32(sm + 128 * 128);
        t0 = clock64();
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int kk = 0; kk < 4; kk++) umma_i8(tmem + (uint32_t)((it & 1) * 256), make_desc_sw128(a0 + kk * 32), make_desc_sw128(b0 + kk * 32), idesc, 1u);
        }
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    if (tid == 0) { t1 = clock64(); cycles_out[blockIdx.x] = t1 - t0; }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------ part C
// Each CTA streams `nsteps` k-slabs: per slab TMA loads SA slices of A rows (128 x 128 B each) and SB slices of B rows (N x 128 B),
// then issues `pairs` x 4 MMAs on them (round-robin over the loaded slices).  2-stage ring.
template <int NST>
__global__ void __launch_bounds__(128) probe_stream(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, int N, int SA, int SB, int pairs,
                                                    int nsteps, int rows_total, unsigned long long* cycles_out) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t full[NST], empty[NST], done;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    const size_t stageA = (size_t)SA * 128 * 128, stageB = (size_t)SB * N * 128, stage = stageA + stageB;
    if (warp == 0) tmem_alloc(&tmem_base, 512);
    if (tid == 0) { for (int s = 0; s < NST; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); } mbar_init(&done, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base;
    unsigned long long t0 = clock64();
    if (warp == 0 && (tid & 31) == 0) {            // TMA producer
        for (int it = 0; it < nsteps; it++) {
            const int s = it % NST;
            if (it >= NST) mbar_wait(&empty[s], ((it / NST) - 1) & 1);
            mbar_expect_tx(&full[s], (uint32_t)stage);
            uint8_t* base = sm + (size_t)s * stage;
            const int row0 = (int)(((size_t)blockIdx.x * 977 + (size_t)it * 131) % (size_t)(rows_total / 256)) * 256;      // wander over the tensor
            for (int a = 0; a < SA; a++) tma_load_2d(base + (size_t)a * 128 * 128, &mapA, (it & 3) * 128, row0 + 0, &full[s]);
            for (int b2 = 0; b2 < SB; b2++) tma_load_2d(base + stageA + (size_t)b2 * N * 128, &mapB, (it & 3) * 128, row0, &full[s]);
        }
    } else if (warp == 1 && (tid & 31) == 0) {     // MMA issuer
        const uint32_t idesc = make_idesc_i8(128, N, 1, 0);
        for (int it = 0; it < nsteps; it++) {
            const int s = it % NST;
            mbar_wait(&full[s], (it / NST) & 1);
            tc_fence_after();
            const uint32_t a0 = smem_u32(sm + (size_t)s * stage), b0 = a0 + (uint32_t)stageA;
            for (int pq = 0; pq < pairs; pq++) {
                const uint32_t aa = a0 + (uint32_t)(pq % SA) * 128 * 128, bb = b0 + (uint32_t)(pq % SB) * (uint32_t)(N * 128);
#pragma unroll
                for (int kk = 0; kk < 4; kk++) umma_i8(tmem + (uint32_t)((pq % (512 / N)) * N), make_desc_sw128(aa + kk * 32), make_desc_sw128(bb + kk * 32), idesc, 1u);
            }
            umma_commit(&empty[s]);
        }
        umma_commit(&done);
    }
    mbar_wait(&done, 0);
    if (tid == 0) cycles_out[blockIdx.x] = clock64() - t0;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make_map(EncodeTiled enc, void* ptr, uint64_t rows, uint64_t kbytes, uint32_t box_rows) {
    CUtensorMap m;
    cuuint64_t dims[2] = {kbytes, rows};
    cuuint64_t strides[1] = {kbytes};
    cuuint32_t box[2] = {128, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); exit(1); }
    return m;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    printf("device: %s, %d SMs, clock %.0f MHz\n", prop.name, prop.multiProcessorCount, prop.clockRate / 1e3);
    const double ghz = prop.clockRate / 1e6;
    // ---------------- part A
    int all_ok = 1;
    for (int cfg = 0; cfg < 4; cfg++) {
        const int N = cfg == 0 ? 128 : (cfg == 1 ? 64 : (cfg == 2 ? 256 : 128)), K = cfg == 3 ? 512 : 256, b_signed = cfg & 1;
        std::vector<int8_t> hA(128 * K); std::vector<uint8_t> hB((size_t)N * K);
        srand(1234 + cfg);
        for (auto& v : hA) v = (int8_t)(rand() % 256 - 128);
        for (auto& v : hB) v = (uint8_t)(rand() % 256);
        int8_t* dA; uint8_t* dB; int32_t* dD;
        CK(cudaMalloc(&dA, hA.size())); CK(cudaMalloc(&dB, hB.size())); CK(cudaMalloc(&dD, (size_t)128 * N * 4));
        CK(cudaMemcpy(dA, hA.data(), hA.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, hB.data(), hB.size(), cudaMemcpyHostToDevice));
        const size_t smem = (size_t)(K / 128) * (128 + N) * 128 + 1024;
        CK(cudaFuncSetAttribute(probe_correct, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        probe_correct<<<1, 128, smem>>>(dA, dB, dD, N, K, b_signed, 1);
        CK(cudaDeviceSynchronize());
        std::vector<int32_t> hD((size_t)128 * N);
        CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
        long long bad = 0;
        for (int i = 0; i < 128; i++) for (int j = 0; j < N; j++) {
            long long acc = 0;
            for (int k = 0; k < K; k++) acc += (long long)hA[(size_t)i * K + k] * (b_signed ? (long long)(int8_t)hB[(size_t)j * K + k] : (long long)hB[(size_t)j * K + k]);
            if ((long long)hD[(size_t)i * N + j] != acc) { if (bad < 4) printf("  mismatch (%d,%d): got %d want %lld\n", i, j, hD[(size_t)i * N + j], acc); bad++; }
        }
        printf("part A: M=128 N=%d K=%d A=s8 B=%s : %s (%lld mismatches)\n", N, K, b_signed ? "s8" : "u8", bad ? "FAIL" : "exact", bad);
        if (bad) all_ok = 0;
        cudaFree(dA); cudaFree(dB); cudaFree(dD);
    }
    if (!all_ok) { printf("part A failed; skipping throughput parts\n"); return 2; }
    EncodeTiled enc = nullptr;
    {
        void* fn = nullptr; cudaDriverEntryPointQueryResult qres;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        enc = (EncodeTiled)fn;
        if (!enc) { printf("no cuTensorMapEncodeTiled\n"); return 3; }
    }
    {   // ---------------- part A2
        constexpr int S = 8;
        const int n = 300, K = 200, Kp = 208;           // ragged sizes: partial row tiles and a partial k-step are zero-filled by TMA
        std::vector<int8_t> h((size_t)S * n * Kp, 0);
        srand(99);
        for (int a = 0; a < S; a++) for (int r = 0; r < n; r++) for (int k = 0; k < K; k++) h[((size_t)a * n + r) * Kp + k] = (int8_t)(rand() % 256 - 128);
        int8_t* dS; int32_t* dD;
        CK(cudaMalloc(&dS, h.size())); CK(cudaMalloc(&dD, (size_t)S * 128 * 64 * 4));
        CK(cudaMemcpy(dS, h.data(), h.size(), cudaMemcpyHostToDevice));
        auto map3 = [&](uint32_t box_rows) {
            CUtensorMap m;
            cuuint64_t dims[3] = {(cuuint64_t)Kp, (cuuint64_t)n, (cuuint64_t)S};
            cuuint64_t strides[2] = {(cuuint64_t)Kp, (cuuint64_t)n * Kp};
            cuuint32_t box[3] = {64, box_rows, (cuuint32_t)S};
            cuuint32_t estr[3] = {1, 1, 1};
            CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, dS, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled(3d) failed: %d\n", (int)r); exit(1); }
            return m;
        };
        CUtensorMap mA = map3(128), mB = map3(64);
        const size_t smem = 2 * (size_t)(S * 128 * 64 + S * 64 * 64) + 1024;
        CK(cudaFuncSetAttribute(probe_sliced<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int rowA0 = 256, rowB0 = 192;             // A tile rows 256..383 (44 valid), B tile rows 192..255
        probe_sliced<S><<<1, 128, smem>>>(mA, mB, dD, K, rowA0, rowB0);
        CK(cudaDeviceSynchronize());
        std::vector<int32_t> hD((size_t)S * 128 * 64);
        CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
        long long bad = 0;
        for (int g = 0; g < S; g++) for (int i = 0; i < 128; i++) for (int j = 0; j < 64; j++) {
            long long acc = 0;
            const int ri = rowA0 + i, rj = rowB0 + j;
            if (ri < n && rj < n)
                for (int a = 0; a <= g; a++) for (int k = 0; k < K; k++) acc += (long long)h[((size_t)a * n + ri) * Kp + k] * (long long)h[((size_t)(g - a) * n + rj) * Kp + k];
            if ((long long)hD[((size_t)g * 128 + i) * 64 + j] != acc) { if (bad < 4) printf("  mismatch g=%d (%d,%d): got %d want %lld\n", g, i, j, hD[((size_t)g * 128 + i) * 64 + j], acc); bad++; }
        }
        printf("part A2: 8 slices, 3-D TMA box + SWIZZLE_64B, 36 slice pairs into 8 TMEM accumulators, ragged n=%d K=%d: %s (%lld mismatches)\n", n, K, bad ? "FAIL" : "exact", bad);
        cudaFree(dS); cudaFree(dD);
        if (bad) return 2;
    }
    // ---------------- part B
    unsigned long long* dcyc; CK(cudaMalloc(&dcyc, 8 * 1024));
    std::vector<unsigned long long> hc(1024);
    const int nsm = prop.multiProcessorCount;
    for (int N : {64, 128, 256}) {
        const int iters = 4096;
        const size_t smem = (128 + 256) * 128 + 1024;
        CK(cudaFuncSetAttribute(probe_issue, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        probe_issue<<<nsm, 128, smem>>>(N, iters, dcyc);
        CK(cudaDeviceSynchronize());
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        CK(cudaEventRecord(e0));
        probe_issue<<<nsm, 128, smem>>>(N, iters, dcyc);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        CK(cudaMemcpy(hc.data(), dcyc, nsm * 8, cudaMemcpyDeviceToHost));
        double avg = 0; for (int i = 0; i < nsm; i++) avg += (double)hc[i]; avg /= nsm;
        const double ops = 2.0 * 128 * N * 32 * 4.0 * iters;
        printf("part B: resident operands, M=128 N=%3d: %.1f cycles per K=32 MMA, %.2f TOP/s per SM (in-kernel clock), %.1f TOP/s whole GPU by events (%.3f ms)\n",
               N, avg / (4.0 * iters), ops / (avg / ghz) * 1e-3, ops * nsm / (ms * 1e-3) * 1e-12, ms);
    }
    // ---------------- part C
    const uint64_t rows_total = 1 << 17, kbytes = 512;       // 64 MiB per operand tensor: L2-resident after the first touch
    uint8_t *gA, *gB;
    CK(cudaMalloc(&gA, rows_total * kbytes)); CK(cudaMalloc(&gB, rows_total * kbytes));
    CK(cudaMemset(gA, 1, rows_total * kbytes)); CK(cudaMemset(gB, 2, rows_total * kbytes));
    struct Cfg { int N, SA, SB, pairs, nst; const char* what; };
    const Cfg cfgs[] = {{128, 1, 1, 1, 2, "plain int8 GEMM tile 128x128, 32 KB stages, 2 stages"},
                        {128, 1, 1, 1, 4, "plain int8 GEMM tile 128x128, 32 KB stages, 4 stages"},
                        {128, 1, 1, 1, 6, "plain int8 GEMM tile 128x128, 32 KB stages, 6 stages"},
                        {64, 4, 4, 10, 2, "Ozaki half: 4+4 planes, 10 pairs, 128x64, 96 KB stages, 2 stages"},
                        {64, 2, 2, 3, 4, "2+2 planes, 3 pairs, 128x64, 48 KB stages, 4 stages"}};
    for (const Cfg& c : cfgs) {
        CUtensorMap mA = make_map(enc, gA, rows_total, kbytes, 128), mB = make_map(enc, gB, rows_total, kbytes, (uint32_t)c.N);
        const size_t stage = (size_t)c.SA * 128 * 128 + (size_t)c.SB * c.N * 128, smem = c.nst * stage + 1024;
        if (smem > 227 * 1024) { printf("part C: %s: %zu B of smem do not fit, skipped\n", c.what, smem); continue; }
        const int nsteps = 2048;
        auto launch = [&]() {
            if (c.nst == 2) { CK(cudaFuncSetAttribute(probe_stream<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); probe_stream<2><<<nsm, 128, smem>>>(mA, mB, c.N, c.SA, c.SB, c.pairs, nsteps, (int)rows_total, dcyc); }
            else if (c.nst == 4) { CK(cudaFuncSetAttribute(probe_stream<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); probe_stream<4><<<nsm, 128, smem>>>(mA, mB, c.N, c.SA, c.SB, c.pairs, nsteps, (int)rows_total, dcyc); }
            else { CK(cudaFuncSetAttribute(probe_stream<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); probe_stream<6><<<nsm, 128, smem>>>(mA, mB, c.N, c.SA, c.SB, c.pairs, nsteps, (int)rows_total, dcyc); }
        };
        launch();
        CK(cudaDeviceSynchronize());
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        const double ops = 2.0 * 128 * c.N * 128.0 * c.pairs * nsteps * nsm, bytes = (double)stage * nsteps * nsm;
        printf("part C: %-68s: %.1f TOP/s, operand stream %.2f TB/s (%.3f ms)\n", c.what, ops / (ms * 1e-3) * 1e-12, bytes / (ms * 1e-3) * 1e-12, ms);
    }
    printf("done\n");
    return 0;
}
