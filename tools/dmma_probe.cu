// tools/dmma_probe.cu -- microbenchmark + layout check of the fp64 mma.sync shapes on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_probe dmma_probe.cu ; run on the GPU box.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma884(double (&d)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ void mma1684(double (&d)[4], const double (&a)[2], double b) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void mma1688(double (&d)[4], const double (&a)[4], const double (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void mma16816(double (&d)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

// ---- layout check: A is 16 x K row-major in global, B is K x 8 (stored as B[k*8+n]), C = A*B (16 x 8)
template <int K>
__global__ void layout_kernel(const double* A, const double* B, double* C) {
    const int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
    double d[4] = {0, 0, 0, 0};
    if (K == 4) { double a[2] = {A[g * K + t], A[(g + 8) * K + t]}; mma1684(d, a, B[t * 8 + g]); }
    if (K == 8) { double a[4], b[2];
        for (int i = 0; i < 4; i++) a[i] = A[(g + 8 * (i & 1)) * K + t + 4 * (i >> 1)];
        for (int i = 0; i < 2; i++) b[i] = B[(t + 4 * i) * 8 + g];
        mma1688(d, a, b); }
    if (K == 16) { double a[8], b[4];
        for (int i = 0; i < 8; i++) a[i] = A[(g + 8 * (i & 1)) * K + t + 4 * (i >> 1)];
        for (int i = 0; i < 4; i++) b[i] = B[(t + 4 * i) * 8 + g];
        mma16816(d, a, b); }
    C[g * 8 + 2 * t] = d[0]; C[g * 8 + 2 * t + 1] = d[1]; C[(g + 8) * 8 + 2 * t] = d[2]; C[(g + 8) * 8 + 2 * t + 1] = d[3];
}
template <int K>
void check_layout() {
    std::vector<double> A(16 * K), B(K * 8), C(128), R(128, 0.0);
    for (auto& v : A) v = rand() % 17 - 8;
    for (auto& v : B) v = rand() % 13 - 6;
    for (int i = 0; i < 16; i++) for (int j = 0; j < 8; j++) for (int k = 0; k < K; k++) R[i * 8 + j] += A[i * K + k] * B[k * 8 + j];
    double *dA, *dB, *dC;
    cudaMalloc(&dA, A.size() * 8); cudaMalloc(&dB, B.size() * 8); cudaMalloc(&dC, 128 * 8);
    cudaMemcpy(dA, A.data(), A.size() * 8, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 8, cudaMemcpyHostToDevice);
    layout_kernel<K><<<1, 32>>>(dA, dB, dC);
    cudaMemcpy(C.data(), dC, 128 * 8, cudaMemcpyDeviceToHost);
    double err = 0; for (int i = 0; i < 128; i++) err = fmax(err, fabs(C[i] - R[i]));
    printf("layout m16n8k%-2d : max err %g  (%s) [%s]\n", K, err, err == 0 ? "OK" : "MISMATCH", cudaGetErrorString(cudaGetLastError()));
}

// ---- throughput: NACC independent accumulator chains per warp, ITER back-to-back rounds
template <int SHAPE, int NACC>
__global__ void tput_kernel(double* out, int iters, double seed) {
    double a2[2] = {seed, seed * 0.5}, a4[4] = {seed, 1, 2, 3}, a8[8] = {seed, 1, 2, 3, 4, 5, 6, 7}, b2[2] = {1, seed}, b4[4] = {1, 2, seed, 4};
    double d2[NACC][2], d4[NACC][4];
    for (int i = 0; i < NACC; i++) { d2[i][0] = d2[i][1] = 0; for (int j = 0; j < 4; j++) d4[i][j] = 0; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) {
            if (SHAPE == 0) mma884(d2[i], a2[0], b2[0]);
            if (SHAPE == 1) mma1684(d4[i], a2, b2[0]);
            if (SHAPE == 2) mma1688(d4[i], a4, b2);
            if (SHAPE == 3) mma16816(d4[i], a8, b4);
        }
    }
    double s = 0;
    for (int i = 0; i < NACC; i++) s += d2[i][0] + d2[i][1] + d4[i][0] + d4[i][1] + d4[i][2] + d4[i][3];
    if (s == 123.456) out[0] = s;
}
template <int SHAPE, int NACC>
void tput(const char* name, double flop_per_mma, int warps_per_sm) {
    double* out; cudaMalloc(&out, 8);
    int sms = 148, iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    tput_kernel<SHAPE, NACC><<<sms, warps_per_sm * 32>>>(out, 100, 1.0);
    cudaEventRecord(e0);
    tput_kernel<SHAPE, NACC><<<sms, warps_per_sm * 32>>>(out, iters, 1.0);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double flops = (double)sms * warps_per_sm * iters * NACC * flop_per_mma;
    printf("tput %-10s nacc=%2d warps/SM=%2d : %8.2f TFLOP/s  (%.3f ms)\n", name, NACC, warps_per_sm, flops / (ms * 1e-3) * 1e-12, ms);
}

__global__ void dfma_kernel(double* out, int iters, double s) {
    double a[16];
    for (int i = 0; i < 16; i++) a[i] = i * s;
    for (int it = 0; it < iters; it++)
#pragma unroll
        for (int i = 0; i < 16; i++) a[i] = fma(a[i], s, 1.0);
    double r = 0; for (int i = 0; i < 16; i++) r += a[i];
    if (r == 1.2345) out[0] = r;
}

int main() {
    check_layout<4>(); check_layout<8>(); check_layout<16>();
    for (int w : {4, 8, 16}) {
        tput<0, 8>("m8n8k4", 2.0 * 8 * 8 * 4, w);
        tput<1, 8>("m16n8k4", 2.0 * 16 * 8 * 4, w);
        tput<2, 8>("m16n8k8", 2.0 * 16 * 8 * 8, w);
        tput<3, 8>("m16n8k16", 2.0 * 16 * 8 * 16, w);
    }
    tput<0, 32>("m8n8k4", 2.0 * 8 * 8 * 4, 8);
    tput<3, 16>("m16n8k16", 2.0 * 16 * 8 * 16, 8);
    {
        double* out; cudaMalloc(&out, 8);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        dfma_kernel<<<148 * 4, 256>>>(out, 100, 1.0000001);
        cudaEventRecord(e0); dfma_kernel<<<148 * 4, 256>>>(out, 20000, 1.0000001); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("tput DFMA (vector fp64): %8.2f TFLOP/s\n", 148.0 * 4 * 256 * 20000 * 16 * 2 / (ms * 1e-3) * 1e-12);
    }
    return 0;
}
