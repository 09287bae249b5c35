// piqp_b200/csrc/dense_backend.cu -- batched dense KKT backend + dense problem data / Ruiz sweeps.
#include "dense_backend.hpp"
#include <string>
#include <cstdlib>
#include <cstdint>
#include <algorithm>
#include <mutex>
#include <vector>
#include <cstdio>
#include "dense_kernels.cuh"

namespace b200 {

std::atomic<unsigned long long> g_launches{0};

// ---- device timeline (common.cuh)
bool g_timeline_on = getenv("B200_TIMELINE") != nullptr;
namespace {
constexpr unsigned long long TIMELINE_CAP = 1ull << 20;
std::mutex tl_mutex;
unsigned long long* tl_buf = nullptr;          // [0] = number of stamps so far; then (time, name id) pairs in execution order
std::vector<const char*> tl_names;
__global__ void timeline_stamp_kernel(unsigned long long* buf, unsigned long long id, unsigned long long cap) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    const unsigned long long slot = atomicAdd(buf, 1ull);
    if (slot < cap) { buf[2 + 2 * slot] = t; buf[3 + 2 * slot] = id; }
}
}  // namespace
void timeline_stamp(const char* name, cudaStream_t stream) {      // the id travels as a kernel argument, so a stamp captured into a graph keeps its name at every replay
    std::lock_guard<std::mutex> lk(tl_mutex);
    if (!tl_buf) { B200_CUDA(cudaMalloc(&tl_buf, 16 * (TIMELINE_CAP + 1))); B200_CUDA(cudaMemset(tl_buf, 0, 16 * (TIMELINE_CAP + 1))); }
    size_t id = 0;
    while (id < tl_names.size() && tl_names[id] != name) id++;     // string literals of one B200_LAUNCH site share an address
    if (id == tl_names.size()) tl_names.push_back(name);
    timeline_stamp_kernel<<<1, 1, 0, stream>>>(tl_buf, id, TIMELINE_CAP);
}
int timeline_dump(const char* path) {
    std::lock_guard<std::mutex> lk(tl_mutex);
    if (!tl_buf) return 0;
    B200_CUDA(device_synchronize_shared());
    unsigned long long cnt = 0;
    B200_CUDA(cudaMemcpy(&cnt, tl_buf, sizeof cnt, cudaMemcpyDeviceToHost));
    cnt = std::min(cnt, TIMELINE_CAP);
    std::vector<unsigned long long> h(2 * cnt);
    if (cnt) B200_CUDA(cudaMemcpy(h.data(), tl_buf + 2, 16 * cnt, cudaMemcpyDeviceToHost));
    FILE* f = fopen(path, "w");
    if (!f) return -1;
    for (unsigned long long i = 0; i < cnt; i++) fprintf(f, "%llu\t%s\t%llu\n", i, h[2 * i + 1] < tl_names.size() ? tl_names[h[2 * i + 1]] : "?", h[2 * i]);
    fclose(f);
    B200_CUDA(cudaMemset(tl_buf, 0, sizeof cnt));      // start over
    return (int)cnt;
}

void BatchedKKT::tic(int kind) {
    if (!profile) return;
    cudaEvent_t e;
    if (free_events_.empty()) B200_CUDA(cudaEventCreate(&e)); else { e = free_events_.back(); free_events_.pop_back(); }
    B200_CUDA(cudaEventRecord(e, stream));
    open_[kind] = e;
}
void BatchedKKT::toc(int kind) {
    if (!profile || !open_[kind]) return;
    cudaEvent_t e;
    if (free_events_.empty()) B200_CUDA(cudaEventCreate(&e)); else { e = free_events_.back(); free_events_.pop_back(); }
    B200_CUDA(cudaEventRecord(e, stream));
    spans_.push_back({open_[kind], e, kind});
    open_[kind] = nullptr;
}
void BatchedKKT::collect() {
    for (auto& s : spans_) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, s.a, s.b) == cudaSuccess) { prof_ms[s.kind] += ms; prof_calls[s.kind]++; }
        free_events_.push_back(s.a); free_events_.push_back(s.b);
    }
    spans_.clear();
}
void BatchedKKT::reset_profile() { collect(); for (int i = 0; i < T_COUNT; i++) { prof_ms[i] = 0; prof_calls[i] = 0; } }

// =====================================================================================================
// data packing
// =====================================================================================================
void DenseData::alloc(int batch_, int n_, int p_, int m_, cudaStream_t zero_stream) {
    batch = batch_; n = n_; p = p_; m = m_; ld = round_up(n > 0 ? n : 1, 8);
    Pf.alloc((size_t)batch * ld * n); AT.alloc((size_t)batch * ld * p); GT.alloc((size_t)batch * ld * m);
    Pf.zero(zero_stream); AT.zero(zero_stream); GT.zero(zero_stream);
}

__global__ void pack_sym_upper_kernel(const double* src, long long sb, long long rs, long long cs, double* Pf, long long sP, int ld, int n) {
    const int b = blockIdx.z;
    const int r = blockIdx.x * blockDim.x + threadIdx.x, c = blockIdx.y;
    if (r >= n) return;
    const int i = min(r, c), j = max(r, c);
    Pf[(size_t)b * sP + (size_t)c * ld + r] = src[(size_t)b * sb + (size_t)i * rs + (size_t)j * cs];
}
void dense_pack_sym_upper(const double* src, long long sb, long long rs, long long cs, DenseData& D, cudaStream_t st) {
    if (D.n == 0) return;
    dim3 grid(ceil_div(D.n, 128), D.n, D.batch);
    B200_LAUNCH(pack_sym_upper_kernel, grid, 128, 0, st, src, sb, rs, cs, D.Pf.get(), D.sP(), D.ld, D.n);
}

__global__ void pack_cols_kernel(const double* src, long long sb, int rows, int ld_src, double* dst, long long sdst, int ld) {
    const int b = blockIdx.z, c = blockIdx.y;
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    dst[(size_t)b * sdst + (size_t)c * ld + r] = src[(size_t)b * sb + (size_t)c * ld_src + r];
}
void dense_pack_cols(const double* src, long long sb, int rows, int cols, int ld_src, double* dst, long long sdst, int ld, int batch, cudaStream_t st) {
    if (rows == 0 || cols == 0) return;
    dim3 grid(ceil_div(rows, 128), cols, batch);
    B200_LAUNCH(pack_cols_kernel, grid, 128, 0, st, src, sb, rows, ld_src, dst, sdst, ld);
}

__global__ void zero_G_rows_kernel(double* GT, long long sG, int ld, int n, int m, const int* mask) {
    const int b = blockIdx.z, k = blockIdx.y;
    if (!mask[(size_t)b * m + k]) return;
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n) GT[(size_t)b * sG + (size_t)k * ld + r] = 0.0;
}
void dense_zero_G_rows(DenseData& D, const int* row_mask, cudaStream_t st) {
    if (D.m == 0) return;
    dim3 grid(ceil_div(D.n, 128), D.m, D.batch);
    B200_LAUNCH(zero_G_rows_kernel, grid, 128, 0, st, D.GT.get(), D.sG(), D.ld, D.n, D.m, row_mask);
}

// =====================================================================================================
// Ruiz equilibration (dense/preconditioner.hpp:64-251), bit-compatible with the oracle's operation order
// =====================================================================================================
void RuizState::alloc(int batch_, int n_, int p_, int m_) {
    batch = batch_; n = n_; p = p_; m = m_;
    const size_t N = (size_t)n + p + m;
    delta.alloc(batch * N); delta_inv.alloc(batch * N); it.alloc(batch * N);
    delta_b.alloc((size_t)batch * n); delta_b_inv.alloc((size_t)batch * n); itb.alloc((size_t)batch * n);
    c.alloc(batch); c_inv.alloc(batch); done.alloc(batch);
}

__device__ __forceinline__ double ruiz_limit(double d) { return d < 1e-4 ? 1.0 : (d > 1e4 ? 1e4 : d); }

__global__ void fill_kernel(double* p, size_t n, double v) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
void fill_async(double* p, size_t n, double v, cudaStream_t st) {
    if (n) B200_LAUNCH(fill_kernel, (unsigned)((n + 255) / 256), 256, 0, st, p, n, v);
}

// sweep start: done[b] |= !(max|1 - it| > eps)
__global__ void ruiz_begin_kernel(const double* it, const double* itb, int N, int n, int* done, double eps) {
    __shared__ double red[32];
    const int b = blockIdx.x;
    double v[1] = {0.0};
    for (int k = threadIdx.x; k < N; k += blockDim.x) v[0] = fmax(v[0], fabs(1.0 - it[(size_t)b * N + k]));
    for (int k = threadIdx.x; k < n; k += blockDim.x) v[0] = fmax(v[0], fabs(1.0 - itb[(size_t)b * n + k]));
    const int op[1] = {RED_MAX};
    block_reduce<1>(v, op, red);
    if (threadIdx.x == 0 && !(v[0] > eps)) done[b] = 1;
}

__global__ void ruiz_rownorm_kernel(const double* Pf, long long sP, const double* AT, long long sA, const double* GT, long long sG,
                                    int ld, int n, int p, int m, const double* xbs, double* it, double* itb, int N, const int* done) {
    const int b = blockIdx.y;
    if (done[b]) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double v = 0.0;
    const double* P = Pf + (size_t)b * sP;
    for (int c = 0; c < n; c++) v = fmax(v, fabs(P[(size_t)c * ld + i]));
    const double* A = AT + (size_t)b * sA;
    for (int c = 0; c < p; c++) v = fmax(v, fabs(A[(size_t)c * ld + i]));
    const double* G = GT + (size_t)b * sG;
    for (int c = 0; c < m; c++) v = fmax(v, fabs(G[(size_t)c * ld + i]));
    const double xb = xbs[(size_t)b * n + i];
    it[(size_t)b * N + i] = fmax(v, xb);
    itb[(size_t)b * n + i] = xb;
}

__global__ void ruiz_colnorm_kernel(const double* AT, long long sA, const double* GT, long long sG, int ld, int n, int p, int m,
                                    double* it, int N, const int* done) {
    const int b = blockIdx.y;
    if (done[b]) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k = blockIdx.x * (blockDim.x >> 5) + warp;
    if (k >= p + m) return;
    const double* col = k < p ? AT + (size_t)b * sA + (size_t)k * ld : GT + (size_t)b * sG + (size_t)(k - p) * ld;
    double v = 0.0;
    for (int i = lane; i < n; i += 32) v = fmax(v, fabs(col[i]));
    v = warp_max(v);
    if (lane == 0) it[(size_t)b * N + n + k] = v;
}

__global__ void ruiz_finalize_kernel(double* it, double* itb, int N, int n, double* c, double* xbs, double* delta, double* delta_b, const int* done) {
    const int b = blockIdx.y;
    if (done[b]) return;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= N) return;
    const double d = 1.0 / sqrt(ruiz_limit(it[(size_t)b * N + k]));
    it[(size_t)b * N + k] = d;
    delta[(size_t)b * N + k] *= d;
    if (k < n) {
        const double db = 1.0 / sqrt(ruiz_limit(itb[(size_t)b * n + k]));
        itb[(size_t)b * n + k] = db;
        c[(size_t)b * n + k] *= d;
        xbs[(size_t)b * n + k] *= db * d;
        delta_b[(size_t)b * n + k] *= db;
    }
}

// P <- gamma*P then D P D ; AT <- Dx AT Dy ; GT <- Dx GT Dz, with the oracle's multiplication order
__global__ void ruiz_scale_kernel(double* Pf, long long sP, double* AT, long long sA, double* GT, long long sG, int ld, int n, int p, int m,
                                  const double* d, int N, const double* gamma, int gamma_inverse, const int* done) {
    const int b = blockIdx.z;
    if (done && done[b]) return;
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const int c = blockIdx.y;
    const double* dv = d + (size_t)b * N;
    if (c < n) {
        double v = Pf[(size_t)b * sP + (size_t)c * ld + r];
        if (gamma) v *= gamma[b];
        const int i = min(r, c), j = max(r, c);
        v = (v * dv[j]) * dv[i];
        Pf[(size_t)b * sP + (size_t)c * ld + r] = v;
    } else if (c < n + p) {
        double* e = AT + (size_t)b * sA + (size_t)(c - n) * ld + r;
        *e = (dv[r] * *e) * dv[c];
    } else {
        double* e = GT + (size_t)b * sG + (size_t)(c - n - p) * ld + r;
        *e = (dv[r] * *e) * dv[c];
    }
}

__global__ void ruiz_scaleP_kernel(double* Pf, long long sP, int ld, int n, const double* gamma) {
    const int b = blockIdx.z, c = blockIdx.y;
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n) Pf[(size_t)b * sP + (size_t)c * ld + r] *= gamma[b];
}

// cost scaling (dense/preconditioner.hpp:141-162): one CTA per instance (rare path, default off)
__global__ void ruiz_cost_kernel(double* Pf, long long sP, int ld, int n, double* c, double* cscale, const int* done) {
    __shared__ double red[64];
    const int b = blockIdx.x;
    if (done[b]) return;
    double* P = Pf + (size_t)b * sP;
    double v[2] = {0.0, 0.0};
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double r = 0.0;
        for (int cc = 0; cc < n; cc++) r = fmax(r, fabs(P[(size_t)cc * ld + i]));
        v[0] += r;
        v[1] = fmax(v[1], fabs(c[(size_t)b * n + i]));
    }
    const int op[2] = {RED_SUM, RED_MAX};
    block_reduce<2>(v, op, red);
    double gamma = ruiz_limit(v[0] / double(n));
    gamma = ruiz_limit(fmax(gamma, v[1]));
    gamma = 1.0 / gamma;
    for (int cc = 0; cc < n; cc++)
        for (int i = threadIdx.x; i < n; i += blockDim.x) P[(size_t)cc * ld + i] *= gamma;
    for (int i = threadIdx.x; i < n; i += blockDim.x) c[(size_t)b * n + i] *= gamma;
    if (threadIdx.x == 0) cscale[b] *= gamma;
}

__global__ void ruiz_inverse_kernel(const double* delta, const double* delta_b, const double* c, double* delta_inv, double* delta_b_inv, double* c_inv, int N, int n) {
    const int b = blockIdx.y;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < N) delta_inv[(size_t)b * N + k] = 1.0 / delta[(size_t)b * N + k];
    if (k < n) delta_b_inv[(size_t)b * n + k] = 1.0 / delta_b[(size_t)b * n + k];
    if (k == 0) c_inv[b] = 1.0 / c[b];
}

// bounds: b *= d_y ; h *= d_z ; x_l,x_u *= d_b  (and for the reuse/unscale path: c, x_b_scaling)
__global__ void ruiz_vectors_kernel(double* c, double* bvec, double* h_l, double* h_u, double* x_l, double* x_u, double* xbs,
                                    const double* d, const double* db, const double* cs, int n, int p, int m, int N, int scale_c_and_xbs) {
    const int b = blockIdx.y;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const double* dv = d + (size_t)b * N;
    if (k < p) bvec[(size_t)b * p + k] *= dv[n + k];
    if (k < m) { h_l[(size_t)b * m + k] *= dv[n + p + k]; h_u[(size_t)b * m + k] *= dv[n + p + k]; }
    if (k < n) {
        const double dbk = db[(size_t)b * n + k];
        x_l[(size_t)b * n + k] *= dbk; x_u[(size_t)b * n + k] *= dbk;
        if (scale_c_and_xbs) {
            c[(size_t)b * n + k] *= cs[b] * dv[k];
            xbs[(size_t)b * n + k] *= dbk * dv[k];
        }
    }
}

void ruiz_launch_begin(RuizState& R, cudaStream_t st) {
    const int N = R.n + R.p + R.m;
    B200_LAUNCH(ruiz_begin_kernel, R.batch, 256, 0, st, R.it.get(), R.itb.get(), N, R.n, R.done.get(), 1e-3);
}
void ruiz_launch_finalize(RuizState& R, double* c, double* xbs, cudaStream_t st) {
    const int N = R.n + R.p + R.m;
    dim3 gridN(ceil_div(std::max(1, std::max(N, R.n)), 256), R.batch);
    B200_LAUNCH(ruiz_finalize_kernel, gridN, 256, 0, st, R.it.get(), R.itb.get(), N, R.n, c, xbs, R.delta.get(), R.delta_b.get(), R.done.get());
}
void ruiz_launch_inverse(RuizState& R, cudaStream_t st) {
    const int N = R.n + R.p + R.m;
    dim3 gridN(ceil_div(std::max(1, std::max(N, R.n)), 256), R.batch);
    B200_LAUNCH(ruiz_inverse_kernel, gridN, 256, 0, st, R.delta.get(), R.delta_b.get(), R.c.get(), R.delta_inv.get(), R.delta_b_inv.get(), R.c_inv.get(), N, R.n);
}
void ruiz_launch_vectors(RuizState& R, double* c, double* b, double* h_l, double* h_u, double* x_l, double* x_u, double* xbs,
                         const double* d, const double* db, const double* cs, int scale_c_and_xbs, cudaStream_t st) {
    const int N = R.n + R.p + R.m;
    dim3 gridN(ceil_div(std::max(1, std::max(N, R.n)), 256), R.batch);
    B200_LAUNCH(ruiz_vectors_kernel, gridN, 256, 0, st, c, b, h_l, h_u, x_l, x_u, xbs, d, db, cs, R.n, R.p, R.m, N, scale_c_and_xbs);
}
void ruiz_reset(RuizState& R, cudaStream_t st) {
    const size_t N = (size_t)R.n + R.p + R.m, B = R.batch;
    fill_async(R.c.get(), B, 1.0, st); fill_async(R.delta.get(), B * N, 1.0, st); fill_async(R.delta_b.get(), B * R.n, 1.0, st);
    fill_async(R.it.get(), B * N, 0.0, st); fill_async(R.itb.get(), B * R.n, 0.0, st);
    B200_CUDA(cudaMemsetAsync(R.done.get(), 0, sizeof(int) * B, st));
}

void dense_ruiz_scale(DenseData& D, RuizState& R, double* c, double* b, double* h_l, double* h_u, double* x_l, double* x_u,
                      double* xbs, bool reuse_prev, bool scale_cost, int max_iter, cudaStream_t st) {
    const int n = D.n, p = D.p, m = D.m, N = n + p + m, B = D.batch;
    const int maxd = std::max(1, std::max(N, n));
    dim3 gridN(ceil_div(maxd, 256), B);
    dim3 gridMat(ceil_div(std::max(n, 1), 128), N, B);
    if (!reuse_prev) {
        fill_async(R.c.get(), B, 1.0, st); fill_async(R.delta.get(), (size_t)B * N, 1.0, st); fill_async(R.delta_b.get(), (size_t)B * n, 1.0, st);
        fill_async(R.it.get(), (size_t)B * N, 0.0, st); fill_async(R.itb.get(), (size_t)B * n, 0.0, st);
        B200_CUDA(cudaMemsetAsync(R.done.get(), 0, sizeof(int) * B, st));
        for (int iter = 0; iter < max_iter; iter++) {
            B200_LAUNCH(ruiz_begin_kernel, B, 256, 0, st, R.it.get(), R.itb.get(), N, n, R.done.get(), 1e-3);
            if (n > 0) {
                dim3 gr(ceil_div(n, 128), B);
                B200_LAUNCH(ruiz_rownorm_kernel, gr, 128, 0, st, D.Pf.get(), D.sP(), D.AT.get(), D.sA(), D.GT.get(), D.sG(), D.ld, n, p, m, xbs,
                            R.it.get(), R.itb.get(), N, R.done.get());
            }
            if (p + m > 0) {
                dim3 gc(ceil_div(p + m, 8), B);
                B200_LAUNCH(ruiz_colnorm_kernel, gc, 256, 0, st, D.AT.get(), D.sA(), D.GT.get(), D.sG(), D.ld, n, p, m, R.it.get(), N, R.done.get());
            }
            B200_LAUNCH(ruiz_finalize_kernel, gridN, 256, 0, st, R.it.get(), R.itb.get(), N, n, c, xbs, R.delta.get(), R.delta_b.get(), R.done.get());
            if (n > 0)
                B200_LAUNCH(ruiz_scale_kernel, gridMat, 128, 0, st, D.Pf.get(), D.sP(), D.AT.get(), D.sA(), D.GT.get(), D.sG(), D.ld, n, p, m,
                            R.it.get(), N, (const double*)nullptr, 0, R.done.get());
            if (scale_cost) B200_LAUNCH(ruiz_cost_kernel, B, 256, 0, st, D.Pf.get(), D.sP(), D.ld, n, c, R.c.get(), R.done.get());
        }
        B200_LAUNCH(ruiz_inverse_kernel, gridN, 256, 0, st, R.delta.get(), R.delta_b.get(), R.c.get(), R.delta_inv.get(), R.delta_b_inv.get(), R.c_inv.get(), N, n);
        B200_LAUNCH(ruiz_vectors_kernel, gridN, 256, 0, st, c, b, h_l, h_u, x_l, x_u, xbs, R.delta.get(), R.delta_b.get(), R.c.get(), n, p, m, N, 0);
    } else {
        if (n > 0)
            B200_LAUNCH(ruiz_scale_kernel, gridMat, 128, 0, st, D.Pf.get(), D.sP(), D.AT.get(), D.sA(), D.GT.get(), D.sG(), D.ld, n, p, m,
                        R.delta.get(), N, R.c.get(), 0, (const int*)nullptr);
        B200_LAUNCH(ruiz_vectors_kernel, gridN, 256, 0, st, c, b, h_l, h_u, x_l, x_u, xbs, R.delta.get(), R.delta_b.get(), R.c.get(), n, p, m, N, 1);
    }
}

void dense_ruiz_unscale(DenseData& D, RuizState& R, double* c, double* b, double* h_l, double* h_u, double* x_l, double* x_u,
                        double* xbs, cudaStream_t st) {
    const int n = D.n, p = D.p, m = D.m, N = n + p + m, B = D.batch;
    const int maxd = std::max(1, std::max(N, n));
    dim3 gridN(ceil_div(maxd, 256), B);
    dim3 gridMat(ceil_div(std::max(n, 1), 128), N, B);
    if (n > 0)
        B200_LAUNCH(ruiz_scale_kernel, gridMat, 128, 0, st, D.Pf.get(), D.sP(), D.AT.get(), D.sA(), D.GT.get(), D.sG(), D.ld, n, p, m,
                    R.delta_inv.get(), N, R.c_inv.get(), 0, (const int*)nullptr);
    B200_LAUNCH(ruiz_vectors_kernel, gridN, 256, 0, st, c, b, h_l, h_u, x_l, x_u, xbs, R.delta_inv.get(), R.delta_b_inv.get(), R.c_inv.get(), n, p, m, N, 1);
}

// =====================================================================================================
// DenseBatchedKKT
// =====================================================================================================
__global__ void inv_and_copy_kernel(const double* z_reg, double* zinv, size_t nz, const double* delta_in, double* delta_out, int batch) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nz) zinv[i] = 1.0 / z_reg[i];
    if (i < (size_t)batch) delta_out[i] = delta_in[i];
}
__global__ void fail_to_ok_kernel(const int* fail, const int* active, int* ok, int batch) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < batch && (!active || active[b])) ok[b] = fail[b] ? 0 : 1;
}
__global__ void clear_fail_kernel(int* fail, const int* active, int batch) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < batch && (!active || active[b])) fail[b] = 0;
}
__global__ void copy_masked_kernel(const double* src, double* dst, int len, const int* active) {
    const int b = blockIdx.y;
    if (active && !active[b]) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < len) dst[(size_t)b * len + i] = src[(size_t)b * len + i];
}
__global__ void extract_diag_kernel(const double* Pf, long long sP, int ld, int n, double* out) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[(size_t)b * n + i] = Pf[(size_t)b * sP + (size_t)i * ld + i];
}

template <class Kern>
static void set_smem(Kern k, size_t bytes) {
    allow_dynamic_smem(k, (size_t)((int)bytes));
}

DenseBatchedKKT::DenseBatchedKKT(DenseData* data, cudaStream_t st) : D(data) {
    batch = D->batch; n = D->n; p = D->p; m = D->m; stream = st;
    K.alloc((size_t)batch * D->ld * n); K.zero(st);
    zinv.alloc((size_t)batch * m); delta.alloc(batch); fail.alloc(batch); fail.zero(st);
    set_smem(gemm_nt_tile_kernel<EPI_ASSEMBLE, true>, GEMM_SMEM);
    set_smem(gemm_nt_t64_kernel<EPI_ASSEMBLE, true>, T64_SMEM);
    set_smem(gemm_nt_t64_kernel<EPI_SUB, false>, T64_SMEM);
    set_smem(gemm_nt_t64_bulk_kernel<EPI_ASSEMBLE, true>, T64_BULK_SMEM);
    set_smem(gemm_nt_t64_bulk_kernel<EPI_SUB, false>, T64_BULK_SMEM);
    chol_split = !(getenv("B200_CHOL_SPLIT") && getenv("B200_CHOL_SPLIT")[0] == '0');      // 0 = the fused round-1 panel kernel
    if (chol_split && !(getenv("B200_CHOL_AUX") && getenv("B200_CHOL_AUX")[0] == '0')) {
        int lo = 0, hi = 0;
        B200_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        B200_CUDA(cudaStreamCreateWithPriority(&chol_aux, cudaStreamNonBlocking, hi));
        chol_ev.resize(2 * (size_t)ceil_div(std::max(n, 1), TILE) + 2);
        for (auto& e : chol_ev) B200_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        B200_CUDA(cudaStreamCreateWithFlags(&chol_aux2, cudaStreamNonBlocking));
        chol_ev2.resize(chol_ev.size());
        for (auto& e : chol_ev2) B200_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        chol_lookahead = !(getenv("B200_CHOL_LOOKAHEAD") && getenv("B200_CHOL_LOOKAHEAD")[0] == '0');
    }
    gemm_t64 = !(getenv("B200_GEMM_T64") && getenv("B200_GEMM_T64")[0] == '0');      // 128 x 64 tiles, two CTAs per SM (default); 0 = the 128 x 128 kernel
    set_smem(gemm_nt_tile_kernel<EPI_ASSEMBLE, false>, GEMM_SMEM);
    set_smem(gemm_nt_tile_kernel<EPI_SUB, false>, GEMM_SMEM);
    set_smem(gemm_nt_tile_kernel<EPI_STORE, false>, GEMM_SMEM);
    set_smem(chol_diag_kernel, CHOL_DIAG_SMEM);
    set_smem(chol_panel_kernel<true>, CHOL_PANEL_SMEM);
    set_smem(chol_panel_kernel<false>, CHOL_PANEL_SMEM);
    set_smem(chol_solve64_kernel, CHOL_SOLVE64_SMEM);
    chol_solve64 = !(getenv("B200_CHOL_SOLVE64") && getenv("B200_CHOL_SOLVE64")[0] == '0');
    Linv_stride = (long long)ceil_div(std::max(n, 1), TILE) * 4 * LB_SZ;
    Linv.alloc((size_t)batch * Linv_stride); Linv.zero(st);
    set_smem(trsv_kernel, (size_t)(n + 32) * sizeof(double) > 48 * 1024 ? (size_t)(n + 32) * sizeof(double) : 48 * 1024);
    if (p > 0) { AtA.alloc((size_t)batch * D->ld * n); AtA.zero(st); compute_AtA(); }
    // ---- Ozaki / tcgen05 assembly: B200_DENSE_ASSEMBLE = ozaki | dmma | auto (default).  Measured (profiles/r01b_dense_sweep.jsonl):
    //      the tcgen05 kernel is exact to the parity bars; its 128 x 64 tiles (TMEM holds 8 accumulators x 64 columns) are bound by the
    //      L2 -> shared-memory operand stream, so it needs a long contraction to amortise its prologue / epilogue: at m = 512 it takes
    //      6.6 ms per launch against 5.0 ms for the DMMA kernel, at m = 1024 8.7 vs 9.2 ms, at m = 2048 12.0 vs 18.6 ms
    //      (= 46 FP64-equivalent TFLOP/s, above the 37 TFLOP/s FP64 pipe).  auto = ozaki from m >= 1024.
    {
        const char* e = getenv("B200_DENSE_ASSEMBLE");
        const std::string mode = e ? e : "auto";
        ozaki = m > 0 && n > 0 && (mode == "ozaki" || (mode == "auto" && m >= 1024 && n >= 512));
    }
    if (ozaki) {
        typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                        CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        void* fn = nullptr; cudaDriverEntryPointQueryResult qres;
        B200_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn) throw std::runtime_error("cuTensorMapEncodeTiled is not available");
        oz_mp = round_up(m, 16);
        if (const char* e2 = getenv("B200_OZ_KSTEP")) oz_kb = atoi(e2) == 64 ? 64 : 32;
        oz_digits.alloc((size_t)batch * OZ_S * n * oz_mp);
        B200_CUDA(cudaMemsetAsync(oz_digits.get(), 0, oz_digits.n, st));
        oz_ex.alloc((size_t)batch * n); oz_sw.alloc((size_t)batch * m); oz_sc.alloc((size_t)batch * n);
        std::vector<int> tiles;
        const int nI = ceil_div(n, OZ_TM), nJ = ceil_div(n, OZ_TN);
        for (int I = 0; I < nI; I++) for (int J = 0; J < nJ && J * OZ_TN < (I + 1) * OZ_TM; J++) { tiles.push_back(I); tiles.push_back(J); }
        oz_ntiles = (int)tiles.size() / 2;
        oz_tiles.alloc(tiles.size());
        B200_CUDA(cudaMemcpy(oz_tiles.get(), tiles.data(), tiles.size() * sizeof(int), cudaMemcpyHostToDevice));
        auto make = [&](unsigned char* out, cuuint32_t box_rows) {
            cuuint64_t dims[3] = {(cuuint64_t)oz_mp, (cuuint64_t)n, (cuuint64_t)OZ_S * batch};
            cuuint64_t strides[2] = {(cuuint64_t)oz_mp, (cuuint64_t)n * oz_mp};
            cuuint32_t box[3] = {(cuuint32_t)oz_kb, box_rows, (cuuint32_t)OZ_S};
            cuuint32_t estr[3] = {1, 1, 1};
            const CUresult r = ((EncodeTiled)fn)(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, oz_digits.get(), dims, strides, box, estr,
                                                 CU_TENSOR_MAP_INTERLEAVE_NONE, oz_kb == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled failed for the digit planes");
        };
        static_assert(sizeof(CUtensorMap) <= 128, "CUtensorMap size");
        make(oz_mapA, OZ_TM); make(oz_mapB, OZ_TN);
        set_smem(oz_gemm_kernel<64, 2>, OZ_GEMM_SMEM);
        set_smem(oz_gemm_kernel<32, 4>, OZ_GEMM_SMEM);
    }
}

// gemm_nt_t64_bulk_kernel (TMA bulk copies + mbarrier ring, no CTA barrier in the main loop) needs whole k-blocks and 16-byte aligned
// rows (an even number of them in memory); everything else runs the cp.async kernel.  B200_GEMM_BULK=0 switches it off.
static bool bulk_gemm_enabled() { const char* e = getenv("B200_GEMM_BULK"); return !(e && e[0] == '0'); }      // read per call: the tests switch it
static bool bulk_gemm_ok(const GemmArgs& g, bool has_w) {
    return bulk_gemm_enabled() && g.K >= 64 && g.K % KB == 0 && g.rows_valid % 2 == 0 && g.lda % 2 == 0 && g.ldb % 2 == 0 && g.strideA % 2 == 0 && g.strideB % 2 == 0 &&
           (!has_w || g.stridew % 2 == 0) && (reinterpret_cast<uintptr_t>(g.A) % 16 == 0) && (reinterpret_cast<uintptr_t>(g.B) % 16 == 0) &&
           (!has_w || reinterpret_cast<uintptr_t>(g.w) % 16 == 0);
}

void dense_syrk_sub_scaled(const double* A, long long strideA, int lda, const double* w, long long stridew, double* C, long long strideC, int ldc,
                           int n, int K, int batch, const int* active, cudaStream_t st, int tj_start, int tj_end, int ncol) {
    if (n <= 0 || K <= 0 || batch <= 0) return;
    static thread_local int configured_dev = -1;
    int dev = 0;
    B200_CUDA(cudaGetDevice(&dev));
    static const bool use_t64 = !(getenv("B200_GEMM_T64") && getenv("B200_GEMM_T64")[0] == '0');      // 128 x 64 tiles, two CTAs per SM (default)
    if (dev != configured_dev) { set_smem(gemm_nt_tile_kernel<EPI_SUB, true>, GEMM_SMEM); set_smem(gemm_nt_t64_kernel<EPI_SUB, true>, T64_SMEM); set_smem(gemm_nt_t64_bulk_kernel<EPI_SUB, true>, T64_BULK_SMEM); configured_dev = dev; }
    GemmArgs g{};
    g.A = A; g.strideA = strideA; g.lda = lda;
    g.B = A; g.strideB = strideA; g.ldb = lda;
    g.w = w; g.stridew = stridew;
    g.C = C; g.strideC = strideC; g.ldc = ldc;
    g.n = n; g.rows_valid = n + (n & 1); g.K = K; g.nt = ceil_div(n, TILE); g.tj_fixed = -1;
    if (tj_end < 0 || tj_end > g.nt) tj_end = g.nt;
    tj_start = std::max(0, std::min(tj_start, tj_end));
    g.tj_start = tj_start; g.ncol = ncol; g.tiles = 0;
    for (int tj = tj_start; tj < tj_end; tj++) g.tiles += g.nt - tj;
    g.active = active;
    if (g.tiles <= 0) return;
    if (use_t64 && K >= 64 && bulk_gemm_ok(g, true)) {
        g.tiles *= 2;
        B200_LAUNCH((gemm_nt_t64_bulk_kernel<EPI_SUB, true>), (unsigned)((size_t)g.tiles * batch), GEMM_THREADS + BULK_PRODUCER_THREADS, T64_BULK_SMEM, st, g);
        return;
    }
    if (use_t64 && K >= 128) { g.tiles *= 2; B200_LAUNCH((gemm_nt_t64_kernel<EPI_SUB, true>), (unsigned)((size_t)g.tiles * batch), GEMM_THREADS, T64_SMEM, st, g); return; }
    B200_LAUNCH((gemm_nt_tile_kernel<EPI_SUB, true>), (unsigned)((size_t)g.tiles * batch), GEMM_THREADS, GEMM_SMEM, st, g);
}

void DenseBatchedKKT::copy_from(const DenseBatchedKKT& o) {
    auto cp = [&](DevBuf<double>& d, const DevBuf<double>& s) { if (s.n) B200_CUDA(cudaMemcpyAsync(d.get(), s.get(), s.n * sizeof(double), cudaMemcpyDeviceToDevice, stream)); };
    cp(K, o.K); cp(AtA, o.AtA); cp(zinv, o.zinv); cp(delta, o.delta); cp(Linv, o.Linv);
    B200_CUDA(cudaMemcpyAsync(fail.get(), o.fail.get(), sizeof(int) * batch, cudaMemcpyDeviceToDevice, stream));
}

void DenseBatchedKKT::compute_AtA() {   // dense/kkt.hpp:53,68
    if (p == 0 || n == 0) return;
    GemmArgs g{};
    g.A = D->AT.get(); g.strideA = D->sA(); g.lda = D->ld;
    g.B = g.A; g.strideB = g.strideA; g.ldb = g.lda;
    g.C = AtA.get(); g.strideC = D->sP(); g.ldc = D->ld;
    g.n = n; g.rows_valid = D->ld; g.K = p; g.nt = ceil_div(n, TILE); g.tj_fixed = -1; g.tiles = g.nt * (g.nt + 1) / 2;
    B200_LAUNCH((gemm_nt_tile_kernel<EPI_STORE, false>), (unsigned)(g.tiles * batch), GEMM_THREADS, GEMM_SMEM, stream, g);
}

void DenseBatchedKKT::update_data(int options) {   // dense/kkt.hpp:62-71
    B200_ZONE("piqp::KKT::update_data");
    if (options & 2) compute_AtA();
}

void DenseBatchedKKT::assemble_ozaki(const double* x_reg, const int* active) {   // dense/kkt.hpp:140-160 on tcgen05 (dense_ozaki.cuh)
    OzArgs a{};
    a.G = D->GT.get(); a.strideG = D->sG(); a.ldg = D->ld;
    a.w = zinv.get(); a.stridew = m;
    a.n = n; a.m = m; a.mp = oz_mp;
    a.Dg = reinterpret_cast<int8_t*>(oz_digits.get()); a.ex = oz_ex.get(); a.sw = oz_sw.get(); a.sc = oz_sc.get();
    a.C = K.get(); a.strideC = D->sP(); a.ldc = D->ld;
    a.Pf = D->Pf.get(); a.strideP = D->sP();
    a.AtA = p > 0 ? AtA.get() : nullptr; a.strideAtA = D->sP();
    a.xreg = x_reg; a.stridex = n; a.delta = delta.get(); a.active = active;
    a.tile_ij = oz_tiles.get(); a.ntiles = oz_ntiles;
    const size_t tot = (size_t)batch * m;
    B200_LAUNCH(oz_sqrt_kernel, (unsigned)((tot + 255) / 256), 256, 0, stream, zinv.get(), oz_sw.get(), tot);
    { dim3 g(ceil_div(n, 128), batch); B200_LAUNCH(oz_rowscale_kernel, g, 128, 0, stream, a); }
    { dim3 g(ceil_div(n, 32), ceil_div(oz_mp, 128), batch); B200_LAUNCH(oz_split_kernel, g, 256, 0, stream, a); }
    if (oz_kb == 64) B200_LAUNCH((oz_gemm_kernel<64, 2>), (unsigned)((size_t)oz_ntiles * batch), 128, OZ_GEMM_SMEM, stream, *reinterpret_cast<const CUtensorMap*>(oz_mapA),
                                 *reinterpret_cast<const CUtensorMap*>(oz_mapB), a);
    else B200_LAUNCH((oz_gemm_kernel<32, 4>), (unsigned)((size_t)oz_ntiles * batch), 128, OZ_GEMM_SMEM, stream, *reinterpret_cast<const CUtensorMap*>(oz_mapA),
                     *reinterpret_cast<const CUtensorMap*>(oz_mapB), a);
}

void DenseBatchedKKT::assemble(const double* x_reg, const int* active) {   // dense/kkt.hpp:140-160
    if (n == 0) return;
    if (ozaki) { assemble_ozaki(x_reg, active); return; }
    GemmArgs g{};
    g.A = D->GT.get(); g.strideA = D->sG(); g.lda = D->ld;
    g.B = g.A; g.strideB = g.strideA; g.ldb = g.lda;
    g.w = zinv.get(); g.stridew = m;
    g.C = K.get(); g.strideC = D->sP(); g.ldc = D->ld;
    g.n = n; g.rows_valid = D->ld; g.K = m; g.nt = ceil_div(n, TILE); g.tj_fixed = -1; g.tiles = g.nt * (g.nt + 1) / 2;
    g.Pf = D->Pf.get(); g.strideP = D->sP();
    g.AtA = p > 0 ? AtA.get() : nullptr; g.strideAtA = D->sP();
    g.xreg = x_reg; g.stridex = n; g.delta = delta.get(); g.active = active; g.fail = nullptr;
    if (m > 0 && gemm_t64 && bulk_gemm_ok(g, true)) { g.tiles *= 2; B200_LAUNCH((gemm_nt_t64_bulk_kernel<EPI_ASSEMBLE, true>), (unsigned)(g.tiles * batch), GEMM_THREADS + BULK_PRODUCER_THREADS, T64_BULK_SMEM, stream, g); }
    else if (m > 0 && gemm_t64) { g.tiles *= 2; B200_LAUNCH((gemm_nt_t64_kernel<EPI_ASSEMBLE, true>), (unsigned)(g.tiles * batch), GEMM_THREADS, T64_SMEM, stream, g); }
    else if (m > 0) B200_LAUNCH((gemm_nt_tile_kernel<EPI_ASSEMBLE, true>), (unsigned)(g.tiles * batch), GEMM_THREADS, GEMM_SMEM, stream, g);
    else { g.A = D->Pf.get(); g.B = g.A; g.K = 0; g.w = nullptr;
           B200_LAUNCH((gemm_nt_tile_kernel<EPI_ASSEMBLE, false>), (unsigned)(g.tiles * batch), GEMM_THREADS, GEMM_SMEM, stream, g); }
}

void DenseBatchedKKT::cholesky(const int* active) {   // Eigen::LLT<Lower>::compute -- see dense_chol.cuh
    if (n == 0) return;
    const int nt = ceil_div(n, TILE);
    B200_LAUNCH(clear_fail_kernel, ceil_div(batch, 256), 256, 0, stream, fail.get(), active, batch);
    auto update = [&](int jb, int t0, int tiles, cudaStream_t st, int k_lo = 0, int k_len = -1) {
        // left-looking update of (part of) block column jb:  K(:, jb) -= L(:, k_lo : k_lo + k_len) L(jb, same)^T  on the two-CTA-per-SM DMMA tile
        // kernel (default: the whole contraction 0 : j0, j0 = 128 .. n - 128); tiles t0 .. t0 + tiles - 1 of the column (two 64-column halves per 128-row tile)
        GemmArgs g{};
        g.A = K.get() + (size_t)k_lo * D->ld; g.strideA = D->sP(); g.lda = D->ld;
        g.B = g.A; g.strideB = g.strideA; g.ldb = g.lda;
        g.C = K.get(); g.strideC = D->sP(); g.ldc = D->ld;
        g.n = n; g.rows_valid = D->ld; g.K = k_len >= 0 ? k_len : jb * TILE; g.nt = nt; g.tj_fixed = -1; g.tj_start = jb; g.tiles = tiles; g.t0 = t0;
        g.active = active; g.fail = fail.get();
        if (bulk_gemm_ok(g, false)) B200_LAUNCH((gemm_nt_t64_bulk_kernel<EPI_SUB, false>), (unsigned)((size_t)tiles * batch), GEMM_THREADS + BULK_PRODUCER_THREADS, T64_BULK_SMEM, st, g);
        else B200_LAUNCH((gemm_nt_t64_kernel<EPI_SUB, false>), (unsigned)((size_t)tiles * batch), GEMM_THREADS, T64_SMEM, st, g);
    };
    for (int jb = 0; jb < nt; jb++) {
        const int j0 = jb * TILE;
        const int rt = nt - jb - 1;
        if (!chol_split) {
            B200_LAUNCH(chol_diag_kernel, batch, CHOL_THREADS, CHOL_DIAG_SMEM, stream, K.get(), D->sP(), D->ld, n, j0, Linv.get(), Linv_stride, fail.get(), active);
            if (rt > 0) B200_LAUNCH(chol_panel_kernel<true>, (unsigned)(rt * batch), CHOL_THREADS, CHOL_PANEL_SMEM, stream, K.get(), D->sP(), D->ld, n, jb, rt,
                                    Linv.get(), Linv_stride, fail.get(), active);
            continue;
        }
        // The factorisation of the diagonal tile is a latency chain (one busy warp per instance, ~65 us): it runs on an auxiliary
        // high-priority stream BESIDE the update of the tiles below it -- a diag CTA (144 KB, 128 registers) leaves room for one
        // tile-kernel CTA on its SM -- and the panel solve joins both.
        const bool fork = chol_aux && jb > 0 && rt > 0;
        // Look-ahead of the diagonal-tile update.  The two half tiles of the diagonal block are 512 CTAs for a 256-QP batch (1.7 waves) with
        // a contraction that grows to K = n - 128, and they sit alone on the critical path in front of the diag factorisation.  All but the
        // last 128 columns of that contraction are final one block column EARLIER, so that part runs on a second side stream beside block
        // column jb - 1 and only the K = 128 remainder stays in front of the diag kernel.
        const bool la = chol_aux2 && chol_lookahead;
        if (la && jb >= 1 && jb + 1 < nt) {                           // early part of the NEXT diagonal tile: columns 0 : j0 of L are final here
            B200_CUDA(cudaEventRecord(chol_ev2[2 * jb], stream));
            B200_CUDA(cudaStreamWaitEvent(chol_aux2, chol_ev2[2 * jb], 0));
            update(jb + 1, 0, 2, chol_aux2, 0, j0);
            B200_CUDA(cudaEventRecord(chol_ev2[2 * jb + 1], chol_aux2));
        }
        if (la && jb >= 2) {                                          // this diagonal tile: join its early part, then the last 128 columns
            B200_CUDA(cudaStreamWaitEvent(stream, chol_ev2[2 * (jb - 1) + 1], 0));
            update(jb, 0, 2, stream, j0 - TILE, TILE);
        } else if (jb > 0) update(jb, 0, 2, stream);                  // the diagonal tile first
        cudaStream_t ds = stream;
        if (fork) {
            B200_CUDA(cudaEventRecord(chol_ev[2 * jb], stream));
            B200_CUDA(cudaStreamWaitEvent(chol_aux, chol_ev[2 * jb], 0));
            ds = chol_aux;
        }
        B200_LAUNCH(chol_diag_kernel, batch, CHOL_THREADS, CHOL_DIAG_SMEM, ds, K.get(), D->sP(), D->ld, n, j0, Linv.get(), Linv_stride, fail.get(), active);
        if (fork) B200_CUDA(cudaEventRecord(chol_ev[2 * jb + 1], chol_aux));
        if (jb > 0 && rt > 0) update(jb, 2, 2 * rt, stream);          // the tiles below, concurrently with the diag factorisation
        if (fork) B200_CUDA(cudaStreamWaitEvent(stream, chol_ev[2 * jb + 1], 0));
        if (rt > 0 && chol_solve64) B200_LAUNCH(chol_solve64_kernel, (unsigned)(2 * rt * batch), CS_THREADS, CHOL_SOLVE64_SMEM, stream, K.get(), D->sP(), D->ld, n, jb, 2 * rt,
                                                Linv.get(), Linv_stride, fail.get(), active);
        else if (rt > 0) B200_LAUNCH(chol_panel_kernel<false>, (unsigned)(rt * batch), CHOL_THREADS, CHOL_PANEL_SMEM, stream, K.get(), D->sP(), D->ld, n, jb, rt,
                                     Linv.get(), Linv_stride, fail.get(), active);
    }
}

void DenseBatchedKKT::factor(const double* delta_in, const double* x_reg, const double* z_reg, const int* active, int* ok) {
    B200_ZONE("piqp::KKT::update_scalings_and_factor");
    const size_t nz = (size_t)batch * m;
    const size_t tot = std::max(nz, (size_t)batch);
    B200_LAUNCH(inv_and_copy_kernel, (unsigned)((tot + 255) / 256), 256, 0, stream, z_reg, zinv.get(), nz, delta_in, delta.get(), batch);
    tic(T_ASSEMBLE);
    assemble(x_reg, active);
    toc(T_ASSEMBLE);
    tic(T_FACTOR);
    cholesky(active);
    toc(T_FACTOR);
    B200_LAUNCH(fail_to_ok_kernel, ceil_div(batch, 256), 256, 0, stream, fail.get(), active, ok, batch);
}

static GemvArgs gemv_args(const double* M, long long sM, int ld, int rows, int cols, const double* x, long long sx, double* z, long long sz, double alpha, const int* active) {
    GemvArgs a{};
    a.M = M; a.strideM = sM; a.ld = ld; a.rows = rows; a.cols = cols; a.x = x; a.stridex = sx; a.z = z; a.stridez = sz; a.alpha = alpha; a.active = active;
    return a;
}

void DenseBatchedKKT::solve(const double* rx, const double* ry, const double* rz, double* lx, double* ly, double* lz, const int* active) {
    // dense/kkt.hpp:86-105
    if (n == 0) return;
    B200_ZONE("piqp::KKT::solve");
    tic(T_SOLVE);
    dim3 gn(ceil_div(n, 256), batch);
    B200_LAUNCH(copy_masked_kernel, gn, 256, 0, stream, rx, lx, n, active);
    if (m > 0) {
        GemvArgs a = gemv_args(D->GT.get(), D->sG(), D->ld, n, m, rz, m, lx, n, 1.0, active);
        a.s = zinv.get(); a.strides = m; a.accumulate = 1;
        B200_LAUNCH(gemv_n_kernel, dim3(ceil_div(n, GEMV_N_ROWS), batch), 256, 0, stream, a);
    }
    if (p > 0) {
        GemvArgs a = gemv_args(D->AT.get(), D->sA(), D->ld, n, p, ry, p, lx, n, 1.0, active);
        a.alpha_v = delta.get(); a.alpha_v_inverse = 1; a.accumulate = 1;
        B200_LAUNCH(gemv_n_kernel, dim3(ceil_div(n, GEMV_N_ROWS), batch), 256, 0, stream, a);
    }
    B200_LAUNCH(trsv_kernel, batch, TRSV_THREADS, (size_t)(n + 32) * sizeof(double), stream, K.get(), D->sP(), D->ld, n, Linv.get(), Linv_stride, lx, (long long)n, active);
    if (p > 0) {
        GemvArgs a = gemv_args(D->AT.get(), D->sA(), D->ld, n, p, lx, n, ly, p, 1.0, active);
        a.alpha_v = delta.get(); a.alpha_v_inverse = 1; a.sub = ry; a.stridesub = p; a.alpha2 = 1.0;
        dim3 g(ceil_div(p, 8), batch);
        B200_LAUNCH(gemv_t_kernel, g, 256, 0, stream, a);
    }
    if (m > 0) {
        GemvArgs a = gemv_args(D->GT.get(), D->sG(), D->ld, n, m, lx, n, lz, m, 1.0, active);
        a.sub = rz; a.stridesub = m; a.alpha2 = 1.0; a.s = zinv.get(); a.strides = m;
        dim3 g(ceil_div(m, 8), batch);
        B200_LAUNCH(gemv_t_kernel, g, 256, 0, stream, a);
    }
    toc(T_SOLVE);
}

void DenseBatchedKKT::eval_P_x(double alpha, const double* x, double* z, const int* active) {   // dense/kkt.hpp:108-114
    B200_ZONE("piqp::KKT::eval_P_x");
    if (n == 0) return;
    dim3 gn(ceil_div(n, 256), batch);
    GemvArgs a = gemv_args(D->Pf.get(), D->sP(), D->ld, n, n, x, n, z, n, alpha, active);
    B200_LAUNCH(gemv_n_kernel, dim3(ceil_div(n, GEMV_N_ROWS), batch), 256, 0, stream, a);
}
void DenseBatchedKKT::eval_A(double an, double at, const double* xn, const double* xt, double* zn, double* zt, const int* active) {   // :117-123
    if (p > 0) {
        GemvArgs a = gemv_args(D->AT.get(), D->sA(), D->ld, n, p, xn, n, zn, p, an, active);
        dim3 g(ceil_div(p, 8), batch);
        B200_LAUNCH(gemv_t_kernel, g, 256, 0, stream, a);
    }
    if (n > 0) {
        GemvArgs a = gemv_args(D->AT.get(), D->sA(), D->ld, n, p, xt, p, zt, n, at, active);
        dim3 gn(ceil_div(n, 256), batch);
        B200_LAUNCH(gemv_n_kernel, dim3(ceil_div(n, GEMV_N_ROWS), batch), 256, 0, stream, a);
    }
}
void DenseBatchedKKT::eval_G(double an, double at, const double* xn, const double* xt, double* zn, double* zt, const int* active) {   // :126-132
    if (m > 0) {
        GemvArgs a = gemv_args(D->GT.get(), D->sG(), D->ld, n, m, xn, n, zn, m, an, active);
        dim3 g(ceil_div(m, 8), batch);
        B200_LAUNCH(gemv_t_kernel, g, 256, 0, stream, a);
    }
    if (n > 0) {
        GemvArgs a = gemv_args(D->GT.get(), D->sG(), D->ld, n, m, xt, m, zt, n, at, active);
        dim3 gn(ceil_div(n, 256), batch);
        B200_LAUNCH(gemv_n_kernel, dim3(ceil_div(n, GEMV_N_ROWS), batch), 256, 0, stream, a);
    }
}
void DenseBatchedKKT::extract_P_diag(double* P_diag) {
    if (n == 0) return;
    dim3 gn(ceil_div(n, 256), batch);
    B200_LAUNCH(extract_diag_kernel, gn, 256, 0, stream, D->Pf.get(), D->sP(), D->ld, n, P_diag);
}
void DenseBatchedKKT::print_info() const {
    printf("b200 dense backend: batch = %d, n = %d, p = %d, m = %d, tile = %d, DMMA m8n8k4 fp64\n", batch, n, p, m, TILE);
}
// algorithmic work per call and instance
double DenseBatchedKKT::factor_flops() const { return (double)n * n * m + (double)n * n * n / 3.0; }
double DenseBatchedKKT::factor_bytes() const { return 8.0 * ((double)n * m + 0.5 * n * n * (2 + (p > 0 ? 1 : 0))) + 8.0 * (n + m); }
double DenseBatchedKKT::solve_flops() const { return 2.0 * n * n + 4.0 * n * m + 4.0 * n * p; }
double DenseBatchedKKT::solve_bytes() const { return 8.0 * ((double)n * n + 2.0 * n * m + 2.0 * n * p); }

}  // namespace b200
