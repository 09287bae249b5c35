// piqp_b200/csrc/capi.cu -- extern "C" surface of libpiqp_b200.so (see include/piqp_b200.h).
#include "../../include/piqp_b200.h"
#include "dense_backend.hpp"
#include "ip_solver.hpp"
#include "multistage_backend.hpp"
#include "sparse_ldlt_backend.hpp"
#include "sparse_data.hpp"
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>
#include <mutex>

using namespace b200;

static thread_local std::string g_err;
static int fail(int code, const std::string& what) { g_err = what; return code; }
#define B200_TRY(...)                                                  \
    try { __VA_ARGS__; }                                               \
    catch (const CudaError& e) { return fail(B200_E_CUDA, e.what()); } \
    catch (const std::exception& e) { return fail(B200_E_INVALID, e.what()); }

// ---------------------------------------------------------------------------------------------------
// single-instance backend handle
// ---------------------------------------------------------------------------------------------------
struct b200kkt_handle {
    int kind = 0;   // 0 dense, 1 multistage, 2 sparse_ldlt
    int device = 0, n = 0, p = 0, m = 0;
    cudaStream_t stream = nullptr;
    DenseData dd;
    std::unique_ptr<DenseBatchedKKT> dense;
    SparseData sd;
    std::unique_ptr<MultistageBatchedKKT> ms;
    std::unique_ptr<SparseLdltBatchedKKT> ldlt;
    BatchedKKT* be = nullptr;
    // staging
    DevBuf<double> stage_mat;   // raw matrix upload
    DevBuf<double> vx[4], vy[4], vz[4];   // n-, p-, m-sized device vectors
    DevBuf<double> delta;
    DevBuf<int> ok;
    ~b200kkt_handle() { if (stream) cudaStreamDestroy(stream); }
};

namespace {

void upload(double* dst, const double* src, size_t cnt, cudaStream_t st) {
    if (cnt) B200_CUDA(cudaMemcpyAsync(dst, src, cnt * sizeof(double), cudaMemcpyHostToDevice, st));
}
void download(double* dst, const double* src, size_t cnt, cudaStream_t st) {
    if (cnt) B200_CUDA(cudaMemcpyAsync(dst, src, cnt * sizeof(double), cudaMemcpyDeviceToHost, st));
}

void dense_single_upload(b200kkt_handle* h, int options, const double* P, const double* AT, const double* GT) {
    const int n = h->n, p = h->p, m = h->m;
    const size_t big = (size_t)n * std::max(n, std::max(p, m));
    if (h->stage_mat.n < big) h->stage_mat.alloc(std::max<size_t>(big, 1));
    if ((options & B200_KKT_UPDATE_P) && n > 0) {
        upload(h->stage_mat.get(), P, (size_t)n * n, h->stream);
        dense_pack_sym_upper(h->stage_mat.get(), 0, 1, n, h->dd, h->stream);   // column-major upper: (i,j) at i + j*n
        B200_CUDA(cudaStreamSynchronize(h->stream));
    }
    if ((options & B200_KKT_UPDATE_A) && p > 0) {
        upload(h->stage_mat.get(), AT, (size_t)n * p, h->stream);
        dense_pack_cols(h->stage_mat.get(), 0, n, p, n, h->dd.AT.get(), h->dd.sA(), h->dd.ld, 1, h->stream);
        B200_CUDA(cudaStreamSynchronize(h->stream));
    }
    if ((options & B200_KKT_UPDATE_G) && m > 0) {
        upload(h->stage_mat.get(), GT, (size_t)n * m, h->stream);
        dense_pack_cols(h->stage_mat.get(), 0, n, m, n, h->dd.GT.get(), h->dd.sG(), h->dd.ld, 1, h->stream);
        B200_CUDA(cudaStreamSynchronize(h->stream));
    }
}

}  // namespace

extern "C" {

const char* b200_last_error(void) { return g_err.c_str(); }
unsigned long long b200_kernel_launch_count(void) { return g_launches.load(); }
int b200_timeline_dump(const char* path) { try { return b200::timeline_dump(path); } catch (...) { return -1; } }
int b200_device_count(void) { int c = 0; if (cudaGetDeviceCount(&c) != cudaSuccess) return 0; return c; }

int b200kkt_dense_create(b200kkt_handle** out, int n, int p, int m, const double* P_utri, const double* AT, const double* GT, int device) {
    if (!out || n < 0 || p < 0 || m < 0 || (n > 0 && !P_utri) || (p > 0 && !AT) || (m > 0 && !GT)) return fail(B200_E_INVALID, "b200kkt_dense_create: bad arguments");
    B200_TRY(
        B200_CUDA(cudaSetDevice(device));
        auto h = std::make_unique<b200kkt_handle>();
        h->device = device; h->n = n; h->p = p; h->m = m;
        B200_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        h->dd.alloc(1, n, p, m);
        B200_CUDA(device_synchronize_shared());
        dense_single_upload(h.get(), 7, P_utri, AT, GT);
        for (int i = 0; i < 4; i++) { h->vx[i].alloc(std::max(n, 1)); h->vy[i].alloc(std::max(p, 1)); h->vz[i].alloc(std::max(m, 1)); }
        h->delta.alloc(1); h->ok.alloc(1);
        h->dense = std::make_unique<DenseBatchedKKT>(&h->dd, h->stream);
        h->be = h->dense.get();
        B200_CUDA(cudaStreamSynchronize(h->stream));
        *out = h.release();
    )
    return B200_OK;
}

namespace {
int sparse_single_create(b200kkt_handle** out, int kind, int n, int p, int m, const int* Pp, const int* Pi, const double* Px,
                         const int* ATp, const int* ATi, const double* ATx, const int* GTp, const int* GTi, const double* GTx, const int* perm, int device, int mode = 0) {
    *out = nullptr;
    B200_TRY(
        B200_CUDA(cudaSetDevice(device));
        auto h = std::make_unique<b200kkt_handle>();
        h->kind = kind; h->device = device; h->n = n; h->p = p; h->m = m;
        B200_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        SparseData& S = h->sd;
        S.n = n; S.p = p; S.m = m;
        S.P.build(n, n, Pp, Pi);
        S.AT.build(n, p, ATp, ATi);
        S.GT.build(n, m, GTp, GTi);
        S.alloc_values(1);
        if (S.P.nnz) B200_CUDA(cudaMemcpy(S.Px.get(), Px, sizeof(double) * S.P.nnz, cudaMemcpyHostToDevice));
        if (S.AT.nnz) B200_CUDA(cudaMemcpy(S.ATx.get(), ATx, sizeof(double) * S.AT.nnz, cudaMemcpyHostToDevice));
        if (S.GT.nnz) B200_CUDA(cudaMemcpy(S.GTx.get(), GTx, sizeof(double) * S.GT.nnz, cudaMemcpyHostToDevice));
        for (int i = 0; i < 4; i++) { h->vx[i].alloc(std::max(n, 1)); h->vy[i].alloc(std::max(p, 1)); h->vz[i].alloc(std::max(m, 1)); }
        h->delta.alloc(1); h->ok.alloc(1);
        B200_CUDA(device_synchronize_shared());
        if (kind == 1) { h->ms = std::make_unique<MultistageBatchedKKT>(&h->sd, h->stream); h->be = h->ms.get(); }
        else { h->ldlt = std::make_unique<SparseLdltBatchedKKT>(&h->sd, perm, h->stream, mode); h->be = h->ldlt.get(); }
        B200_CUDA(cudaStreamSynchronize(h->stream));
        *out = h.release();
    )
    return B200_OK;
}
}  // namespace

int b200kkt_sparse_create(b200kkt_handle** out, int n, int p, int m, const int* Pp, const int* Pi, const double* Px, const int* ATp, const int* ATi, const double* ATx,
                          const int* GTp, const int* GTi, const double* GTx, int mode, const int* perm, int device) {
    if (!out || n <= 0 || p < 0 || m < 0 || !Pp || (p > 0 && !ATp) || (m > 0 && !GTp)) return fail(B200_E_INVALID, "b200kkt_sparse_create: bad arguments");
    if (mode < 0 || mode > 3) { *out = nullptr; return fail(B200_E_INVALID, "b200kkt_sparse_create: mode must be a KKTMode (0..3)"); }
    return sparse_single_create(out, 2, n, p, m, Pp, Pi, Px, ATp, ATi, ATx, GTp, GTi, GTx, perm, device, mode);
}
int b200kkt_multistage_create(b200kkt_handle** out, int n, int p, int m, const int* Pp, const int* Pi, const double* Px,
                              const int* ATp, const int* ATi, const double* ATx, const int* GTp, const int* GTi, const double* GTx, int device) {
    if (!out || n <= 0 || p < 0 || m < 0 || !Pp || (p > 0 && !ATp) || (m > 0 && !GTp)) return fail(B200_E_INVALID, "b200kkt_multistage_create: bad arguments");
    return sparse_single_create(out, 1, n, p, m, Pp, Pi, Px, ATp, ATi, ATx, GTp, GTi, GTx, nullptr, device);
}
int b200_sparse_ldlt_symbolic(int n, int p, int m, const int* Pp, const int* Pi, const int* ATp, const int* ATi, const int* GTp, const int* GTi,
                              const int* perm_in, int* perm_out, long long* nnz_kkt, long long* nnz_L, int* levels, double* factor_flops) {
    return b200_sparse_ldlt_symbolic_mode(n, p, m, Pp, Pi, ATp, ATi, GTp, GTi, 0, perm_in, perm_out, nnz_kkt, nnz_L, levels, factor_flops, nullptr, nullptr);
}
int b200_sparse_ldlt_symbolic_mode(int n, int p, int m, const int* Pp, const int* Pi, const int* ATp, const int* ATi, const int* GTp, const int* GTi, int mode,
                                   const int* perm_in, int* perm_out, long long* nnz_kkt, long long* nnz_L, int* levels, double* factor_flops,
                                   int* n_supernodes, int* largest_front) {
    if (n <= 0 || p < 0 || m < 0 || !Pp || (p > 0 && !ATp) || (m > 0 && !GTp) || mode < 0 || mode > 3) return fail(B200_E_INVALID, "b200_sparse_ldlt_symbolic: bad arguments");
    B200_TRY(
        auto host_pattern = [](Pattern& M, int rows, int cols, const int* cp, const int* ri) {
            M.rows = rows; M.cols = cols;
            if (cp) M.p.assign(cp, cp + cols + 1); else M.p.assign(cols + 1, 0);
            M.nnz = M.p[cols];
            if (M.nnz) M.i.assign(ri, ri + M.nnz);
        };
        Pattern P, AT, GT;
        host_pattern(P, n, n, Pp, Pi); host_pattern(AT, n, p, ATp, ATi); host_pattern(GT, n, m, GTp, GTi);
        LdltSymbolic S;
        if (!S.analyse(P, AT, GT, perm_in, mode)) throw std::runtime_error(S.error);
        if (n_supernodes) *n_supernodes = S.nsup;
        if (largest_front) *largest_front = S.fmax;
        if (perm_out) std::copy(S.perm.begin(), S.perm.end(), perm_out);
        if (nnz_kkt) *nnz_kkt = (long long)S.Ki.size();      // structural entries (supernode amalgamation pads PK with explicit zeros)
        if (nnz_L) *nnz_L = (long long)S.nnzL();
        if (levels) *levels = (int)S.level_ptr.size() - 1;
        if (factor_flops) *factor_flops = S.factor_flops();
    )
    return B200_OK;
}
int b200kkt_sparse_info(b200kkt_handle* h, long long* nnz_kkt, long long* nnz_L, int* levels, int* perm) {
    if (!h || !h->ldlt) return fail(B200_E_INVALID, "not a sparse_ldlt handle");
    const LdltSymbolic& S = h->ldlt->S;
    if (nnz_kkt) *nnz_kkt = (long long)S.Ki.size();      // structural entries (supernode amalgamation pads PK with explicit zeros)
    if (nnz_L) *nnz_L = (long long)S.nnzL();
    if (levels) *levels = (int)S.level_ptr.size() - 1;
    if (perm) std::copy(S.perm.begin(), S.perm.end(), perm);
    return B200_OK;
}

int b200kkt_update_data(b200kkt_handle* h, int options, const double* P, const double* AT, const double* GT) {
    if (!h) return fail(B200_E_INVALID, "null handle");
    B200_TRY(
        B200_CUDA(cudaSetDevice(h->device));
        if (h->kind == 0) dense_single_upload(h, options, P, AT, GT);
        else {
            SparseData& S = h->sd;
            if ((options & B200_KKT_UPDATE_P) && S.P.nnz) B200_CUDA(cudaMemcpyAsync(S.Px.get(), P, sizeof(double) * S.P.nnz, cudaMemcpyHostToDevice, h->stream));
            if ((options & B200_KKT_UPDATE_A) && S.AT.nnz) B200_CUDA(cudaMemcpyAsync(S.ATx.get(), AT, sizeof(double) * S.AT.nnz, cudaMemcpyHostToDevice, h->stream));
            if ((options & B200_KKT_UPDATE_G) && S.GT.nnz) B200_CUDA(cudaMemcpyAsync(S.GTx.get(), GT, sizeof(double) * S.GT.nnz, cudaMemcpyHostToDevice, h->stream));
        }
        h->be->update_data(options);
        B200_CUDA(cudaStreamSynchronize(h->stream));
    )
    return B200_OK;
}

int b200kkt_factor(b200kkt_handle* h, double delta, const double* x_reg, const double* z_reg) {
    if (!h) return fail(B200_E_INVALID, "null handle");
    int ok = 0;
    B200_TRY(
        B200_CUDA(cudaSetDevice(h->device));
        upload(h->delta.get(), &delta, 1, h->stream);
        // x_reg / z_reg live in their own buffers (vx[3], vz[3]): solve / eval_* stage through vx[0..1] and must not clobber
        // the scalings of the current factorisation (b200kkt_dense_get_kkt re-assembles from them)
        upload(h->vx[3].get(), x_reg, h->n, h->stream);
        upload(h->vz[3].get(), z_reg, h->m, h->stream);
        h->be->factor(h->delta.get(), h->vx[3].get(), h->vz[3].get(), nullptr, h->ok.get());
        B200_CUDA(cudaMemcpyAsync(&ok, h->ok.get(), sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        B200_CUDA(cudaStreamSynchronize(h->stream));
    )
    return ok ? 1 : 0;
}

int b200kkt_solve(b200kkt_handle* h, const double* rx, const double* ry, const double* rz, double* lx, double* ly, double* lz) {
    if (!h) return fail(B200_E_INVALID, "null handle");
    B200_TRY(
        B200_CUDA(cudaSetDevice(h->device));
        upload(h->vx[0].get(), rx, h->n, h->stream); upload(h->vy[0].get(), ry, h->p, h->stream); upload(h->vz[0].get(), rz, h->m, h->stream);
        h->be->solve(h->vx[0].get(), h->vy[0].get(), h->vz[0].get(), h->vx[1].get(), h->vy[1].get(), h->vz[1].get(), nullptr);
        download(lx, h->vx[1].get(), h->n, h->stream); download(ly, h->vy[1].get(), h->p, h->stream); download(lz, h->vz[1].get(), h->m, h->stream);
        B200_CUDA(cudaStreamSynchronize(h->stream));
    )
    return B200_OK;
}

int b200kkt_eval_P_x(b200kkt_handle* h, double alpha, const double* x, double* z) {
    if (!h) return fail(B200_E_INVALID, "null handle");
    B200_TRY(
        B200_CUDA(cudaSetDevice(h->device));
        upload(h->vx[0].get(), x, h->n, h->stream);
        h->be->eval_P_x(alpha, h->vx[0].get(), h->vx[1].get(), nullptr);
        download(z, h->vx[1].get(), h->n, h->stream);
        B200_CUDA(cudaStreamSynchronize(h->stream));
    )
    return B200_OK;
}
int b200kkt_eval_A_xn_and_AT_xt(b200kkt_handle* h, double an, double at, const double* xn, const double* xt, double* zn, double* zt) {
    if (!h) return fail(B200_E_INVALID, "null handle");
    B200_TRY(
        B200_CUDA(cudaSetDevice(h->device));
        upload(h->vx[0].get(), xn, h->n, h->stream); upload(h->vy[0].get(), xt, h->p, h->stream);
        h->be->eval_A(an, at, h->vx[0].get(), h->vy[0].get(), h->vy[1].get(), h->vx[1].get(), nullptr);
        download(zn, h->vy[1].get(), h->p, h->stream); download(zt, h->vx[1].get(), h->n, h->stream);
        B200_CUDA(cudaStreamSynchronize(h->stream));
    )
    return B200_OK;
}
int b200kkt_eval_G_xn_and_GT_xt(b200kkt_handle* h, double an, double at, const double* xn, const double* xt, double* zn, double* zt) {
    if (!h) return fail(B200_E_INVALID, "null handle");
    B200_TRY(
        B200_CUDA(cudaSetDevice(h->device));
        upload(h->vx[0].get(), xn, h->n, h->stream); upload(h->vz[0].get(), xt, h->m, h->stream);
        h->be->eval_G(an, at, h->vx[0].get(), h->vz[0].get(), h->vz[1].get(), h->vx[1].get(), nullptr);
        download(zn, h->vz[1].get(), h->m, h->stream); download(zt, h->vx[1].get(), h->n, h->stream);
        B200_CUDA(cudaStreamSynchronize(h->stream));
    )
    return B200_OK;
}

b200kkt_handle* b200kkt_clone(const b200kkt_handle* src) {
    if (!src) return nullptr;
    try {
        B200_CUDA(cudaSetDevice(src->device));
        auto h = std::make_unique<b200kkt_handle>();
        h->device = src->device; h->n = src->n; h->p = src->p; h->m = src->m; h->kind = src->kind;
        B200_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        if (src->kind != 0) {
            const SparseData& O = src->sd; SparseData& S = h->sd;
            S.n = O.n; S.p = O.p; S.m = O.m;
            S.P.build(O.P.rows, O.P.cols, O.P.p.data(), O.P.i.data()); S.AT.build(O.AT.rows, O.AT.cols, O.AT.p.data(), O.AT.i.data());
            S.GT.build(O.GT.rows, O.GT.cols, O.GT.p.data(), O.GT.i.data());
            S.alloc_values(1);
            auto cpv = [&](DevBuf<double>& d, const DevBuf<double>& s) { if (s.n) B200_CUDA(cudaMemcpy(d.get(), s.get(), s.n * sizeof(double), cudaMemcpyDeviceToDevice)); };
            cpv(S.Px, O.Px); cpv(S.ATx, O.ATx); cpv(S.GTx, O.GTx);
            for (int i = 0; i < 4; i++) { h->vx[i].alloc(std::max(h->n, 1)); h->vy[i].alloc(std::max(h->p, 1)); h->vz[i].alloc(std::max(h->m, 1)); }
            h->delta.alloc(1); h->ok.alloc(1);
            B200_CUDA(device_synchronize_shared());
            if (src->kind == 1) {
                h->ms = std::make_unique<MultistageBatchedKKT>(&h->sd, h->stream);
                h->ms->copy_from(*src->ms);
                h->be = h->ms.get();
            } else {
                h->ldlt = std::make_unique<SparseLdltBatchedKKT>(&h->sd, src->ldlt->S.perm.data(), h->stream, src->ldlt->S.mode);
                h->ldlt->copy_from(*src->ldlt);
                h->be = h->ldlt.get();
            }
            B200_CUDA(cudaStreamSynchronize(h->stream));
            return h.release();
        }
        h->dd.alloc(1, h->n, h->p, h->m);
        B200_CUDA(device_synchronize_shared());
        auto cp = [&](DevBuf<double>& d, const DevBuf<double>& s) { if (s.n) B200_CUDA(cudaMemcpyAsync(d.get(), s.get(), s.n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream)); };
        cp(h->dd.Pf, src->dd.Pf); cp(h->dd.AT, src->dd.AT); cp(h->dd.GT, src->dd.GT);
        for (int i = 0; i < 4; i++) { h->vx[i].alloc(std::max(h->n, 1)); h->vy[i].alloc(std::max(h->p, 1)); h->vz[i].alloc(std::max(h->m, 1)); }
        h->delta.alloc(1); h->ok.alloc(1);
        h->dense = std::make_unique<DenseBatchedKKT>(&h->dd, h->stream);
        h->dense->copy_from(*src->dense);
        h->be = h->dense.get();
        B200_CUDA(cudaStreamSynchronize(h->stream));
        return h.release();
    } catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}
void b200kkt_print_info(const b200kkt_handle* h) { if (h && h->be) h->be->print_info(); }
void b200kkt_destroy(b200kkt_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    ReleaseScope scope;
    delete h;
}

int b200kkt_dense_get_kkt(b200kkt_handle* h, double* kkt_lower, double* chol_lower) {
    if (!h || !h->dense) return fail(B200_E_INVALID, "not a dense handle");
    B200_TRY(
        B200_CUDA(cudaSetDevice(h->device));
        const int n = h->n, ld = h->dd.ld;
        std::vector<double> buf((size_t)ld * n);
        auto fetch = [&](double* dst) {
            download(buf.data(), h->dense->K.get(), buf.size(), h->stream);
            B200_CUDA(cudaStreamSynchronize(h->stream));
            for (int j = 0; j < n; j++) for (int i = 0; i < n; i++) dst[i + (size_t)j * n] = i >= j ? buf[i + (size_t)j * ld] : 0.0;
        };
        if (chol_lower) fetch(chol_lower);
        if (kkt_lower) {   // re-assemble with the scalings of the last factor call (kept in vx[3])
            h->dense->assemble(h->vx[3].get(), nullptr);
            fetch(kkt_lower);
            h->dense->cholesky(nullptr);
            B200_CUDA(cudaStreamSynchronize(h->stream));
        }
    )
    return B200_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------
// batched QP solver handle
// ---------------------------------------------------------------------------------------------------
struct b200qp_handle {
    int device = 0, batch = 0, n = 0, p = 0, m = 0;
    cudaStream_t stream = nullptr;
    b200qp_settings st{};
    int kind = 0;   // 0 dense, 1 sparse pattern + multistage backend
    DenseData dd;
    SparseData sd;
    std::vector<int> A_map, G_map, P_map;    // internal value index -> index in the caller's CSC value arrays
    DevBuf<int> d_A_map, d_G_map, d_P_map;
    int nnzP_in = 0, nnzA_in = 0, nnzG_in = 0;
    RuizState ruiz;
    std::unique_ptr<DenseBatchedKKT> dense;
    std::unique_ptr<MultistageBatchedKKT> ms;
    std::unique_ptr<SparseLdltBatchedKKT> ldlt;
    BatchedKKT* be = nullptr;
    std::unique_ptr<BatchedIPSolver> ip;
    DevBuf<int> zero_rows;
    bool solved = false;
    double setup_ms = 0, update_ms = 0;
    ~b200qp_handle() { ip.reset(); dense.reset(); ms.reset(); ldlt.reset(); if (stream) cudaStreamDestroy(stream); }
};

namespace {

// bounds bookkeeping of Data::set_h_l / set_h_u / disable_inf_constraints / set_x_l / set_x_u (dense/data.hpp:100-206)
__global__ void k_setup_bounds(IpDev d, const double* h_l, const double* h_u, const double* x_l, const double* x_u,
                               int set_hl, int set_hu, int set_xl, int set_xu, int* zero_rows) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < d.m && (set_hl || set_hu)) {
        const size_t k = (size_t)b * d.m + i;
        double lo = d.h_l[k], hi = d.h_u[k];
        int hl = d.has_hl[k], hu = d.has_hu[k];
        if (set_hl) { const double v = h_l ? h_l[k] : -INFINITY; hl = v > -kInf; lo = hl ? v : -kInf; }
        if (set_hu) { const double v = h_u ? h_u[k] : INFINITY; hu = v < kInf; hi = hu ? v : kInf; }
        int zr = 0;
        if (!hl && !hu) { lo = -1.0; hi = 1.0; hl = 1; hu = 1; zr = 1; }
        d.h_l[k] = lo; d.h_u[k] = hi; d.has_hl[k] = hl; d.has_hu[k] = hu; zero_rows[k] = zr;
    }
    if (i < d.n) {
        const size_t k = (size_t)b * d.n + i;
        if (set_xl) { const double v = x_l ? x_l[k] : -INFINITY; const int h = v > -kInf; d.has_xl[k] = h; d.x_l[k] = h ? v : 0.0; }
        if (set_xu) { const double v = x_u ? x_u[k] : INFINITY; const int h = v < kInf; d.has_xu[k] = h; d.x_u[k] = h ? v : 0.0; }
    }
}
__global__ void k_fill_d(double* p, size_t n, double v) { size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = v; }

struct Staged {   // host or device source -> device pointer
    DevBuf<double> buf;
    const double* get(const double* src, size_t cnt, int on_device, cudaStream_t st) {
        if (!src || cnt == 0) return nullptr;
        if (on_device) return src;
        buf.alloc(cnt);
        B200_CUDA(cudaMemcpyAsync(buf.get(), src, cnt * sizeof(double), cudaMemcpyHostToDevice, st));
        return buf.get();
    }
};

void copy_vec(double* dst, const double* src, size_t cnt, int on_device, cudaStream_t st) {
    if (!src || !cnt) return;
    B200_CUDA(cudaMemcpyAsync(dst, src, cnt * sizeof(double), on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
}

void load_problem(b200qp_handle* h, bool first, const double* P, const double* c, const double* A, const double* b, const double* G,
                  const double* h_l, const double* h_u, const double* x_l, const double* x_u, int on_device) {
    const int B = h->batch, n = h->n, p = h->p, m = h->m;
    cudaStream_t st = h->stream;
    IpDev& d = h->ip->dev();
    // Host buffers: the H2D phases of handles that are set up concurrently from several host threads take turns (one process-wide
    // turnstile held until this handle's copies have landed), so each gets the whole PCIe link and the NEXT handle's copies overlap this
    // handle's Ruiz sweeps and solve -- the pipelining of sub-batches that bench.py used to arrange with its own events.
    static std::mutex h2d_turnstile;
    std::unique_lock<std::mutex> turn(h2d_turnstile, std::defer_lock);
    if (!on_device) turn.lock();
    // staging buffers are released in stream order (no device-wide wait): several handles may be set up and solved concurrently
    // from different host threads, and the H2D copies of one then overlap the kernels of the others
    {
        Staged s;
        if (P && n > 0) {
            const double* src = nullptr;
            if (!on_device && n >= 256) {
                // Only the upper triangle of P is read (piqp_typedef.h:44, dense/data.hpp): upload it as trapezoids of 128 rows
                // (columns r0 .. n-1 of rows r0 .. r0+127) -- 56 % of the bytes of the full matrices at n = 1024, and P is two thirds
                // of what config 2 sends over PCIe.  The lower part of the staging buffer stays untouched and is never read.
                s.buf.alloc((size_t)B * n * n);
                for (int b2 = 0; b2 < B; b2++)
                    for (int r0 = 0; r0 < n; r0 += 128) {
                        const int rows = std::min(128, n - r0);
                        const size_t off = (size_t)b2 * n * n + (size_t)r0 * n + r0;
                        B200_CUDA(cudaMemcpy2DAsync(s.buf.get() + off, (size_t)n * sizeof(double), P + off, (size_t)n * sizeof(double), (size_t)(n - r0) * sizeof(double), rows,
                                                    cudaMemcpyHostToDevice, st));
                    }
                src = s.buf.get();
            } else src = s.get(P, (size_t)B * n * n, on_device, st);
            dense_pack_sym_upper(src, (long long)n * n, n, 1, h->dd, st);
            s.buf.release_on(st);
        }
    }
    {
        Staged s;
        if (A && p > 0) { const double* src = s.get(A, (size_t)B * p * n, on_device, st); dense_pack_cols(src, (long long)p * n, n, p, n, h->dd.AT.get(), h->dd.sA(), h->dd.ld, B, st); s.buf.release_on(st); }
    }
    {
        Staged s;
        if (G && m > 0) { const double* src = s.get(G, (size_t)B * m * n, on_device, st); dense_pack_cols(src, (long long)m * n, n, m, n, h->dd.GT.get(), h->dd.sG(), h->dd.ld, B, st); s.buf.release_on(st); }
    }
    copy_vec(d.c, c, (size_t)B * n, on_device, st);
    copy_vec(d.b, b, (size_t)B * p, on_device, st);
    Staged s1, s2, s3, s4;
    const double* dhl = s1.get(h_l, (size_t)B * m, on_device, st);
    const double* dhu = s2.get(h_u, (size_t)B * m, on_device, st);
    const double* dxl = s3.get(x_l, (size_t)B * n, on_device, st);
    const double* dxu = s4.get(x_u, (size_t)B * n, on_device, st);
    const int set_hl = first || h_l, set_hu = first || h_u, set_xl = first || x_l, set_xu = first || x_u;
    const int len = std::max(1, std::max(n, m));
    dim3 grid(ceil_div(len, 256), B);
    B200_LAUNCH(k_setup_bounds, grid, 256, 0, st, d, dhl, dhu, dxl, dxu, set_hl, set_hu, set_xl, set_xu, h->zero_rows.get());
    if (m > 0 && (set_hl || set_hu)) dense_zero_G_rows(h->dd, h->zero_rows.get(), st);
    s1.buf.release_on(st); s2.buf.release_on(st); s3.buf.release_on(st); s4.buf.release_on(st);
    B200_CUDA(cudaStreamSynchronize(st));      // the caller may reuse its host buffers when this returns
}

void fill_d(double* p, size_t n, double v, cudaStream_t st) { if (n) B200_LAUNCH(k_fill_d, (unsigned)((n + 255) / 256), 256, 0, st, p, n, v); }

void scale_problem(b200qp_handle* h, bool reuse) {
    IpDev& d = h->ip->dev();
    RuizState& R = h->ruiz;
    dense_ruiz_scale(h->dd, R, d.c, d.b, d.h_l, d.h_u, d.x_l, d.x_u, d.xbs, reuse, h->st.preconditioner_scale_cost != 0, h->st.preconditioner_iter, h->stream);
}

}  // namespace

extern "C" {

void b200qp_set_default_settings_dense(b200qp_settings* s) {   // settings.hpp:45-82 ; solver.hpp:56-63
    if (!s) return;
    s->rho_init = 1e-6; s->delta_init = 1e-4; s->eps_abs = 1e-8; s->eps_rel = 1e-9; s->check_duality_gap = 1;
    s->eps_duality_gap_abs = 1e-8; s->eps_duality_gap_rel = 1e-9; s->infeasibility_threshold = 0.9;
    s->reg_lower_limit = 1e-10; s->reg_finetune_lower_limit = 1e-13; s->reg_finetune_primal_update_threshold = 7;
    s->reg_finetune_dual_update_threshold = 7; s->max_iter = 250; s->max_factor_retires = 10;
    s->preconditioner_scale_cost = 0; s->preconditioner_reuse_on_update = 0; s->preconditioner_iter = 10; s->tau = 0.99;
    s->kkt_solver = 0; s->iterative_refinement_always_enabled = 0; s->iterative_refinement_eps_abs = 1e-12;
    s->iterative_refinement_eps_rel = 1e-12; s->iterative_refinement_max_iter = 10; s->iterative_refinement_min_improvement_rate = 5.0;
    s->iterative_refinement_static_regularization_eps = 1e-8;
    s->iterative_refinement_static_regularization_rel = 2.220446049250313e-16 * 2.220446049250313e-16;
    s->verbose = 0; s->compute_timings = 0;
}
void b200qp_set_default_settings_sparse(b200qp_settings* s) { b200qp_set_default_settings_dense(s); if (s) s->kkt_solver = 1; }

int b200qp_setup_dense(b200qp_handle** out, int batch, int n, int p, int m, const double* P, const double* c, const double* A, const double* b,
                       const double* G, const double* h_l, const double* h_u, const double* x_l, const double* x_u,
                       const b200qp_settings* settings, int device, int on_device) {
    if (!out || batch <= 0 || n <= 0 || p < 0 || m < 0 || !P || !c) return fail(B200_E_INVALID, "b200qp_setup_dense: bad arguments");
    if ((p > 0 && (!A || !b)) || (m > 0 && (!G || (!h_l && !h_u)))) return fail(B200_E_INVALID, "b200qp_setup_dense: missing constraint data");
    if (batch > B200_MAX_BATCH) return fail(B200_E_UNSUPPORTED, "b200qp_setup_dense: batch > 65535 (the batch index is a gridDim.y / .z coordinate); split the batch over several handles");
    B200_ZONE("piqp::Solver::setup");
    B200_TRY(
        B200_CUDA(cudaSetDevice(device));
        auto h = std::make_unique<b200qp_handle>();
        h->device = device; h->batch = batch; h->n = n; h->p = p; h->m = m;
        if (settings) h->st = *settings; else b200qp_set_default_settings_dense(&h->st);
        B200_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        cudaEvent_t e0, e1;
        B200_CUDA(cudaEventCreate(&e0)); B200_CUDA(cudaEventCreate(&e1));
        const bool verbose_timing = getenv("B200_TIMING") != nullptr;
        auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
        double t_0 = now();
        h->dd.alloc(batch, n, p, m, h->stream);          // zero-fill on the handle's stream: no device-wide synchronisation in the setup of a pipelined sub-batch
        B200_CUDA(cudaEventRecord(e0, h->stream));
        h->ip = std::make_unique<BatchedIPSolver>(batch, n, p, m, h->st, h->stream);
        h->ruiz.alloc(batch, n, p, m);
        h->zero_rows.alloc(std::max<size_t>((size_t)batch * m, 1)); h->zero_rows.zero(h->stream);
        IpDev& d = h->ip->dev();
        fill_d(d.xbs, (size_t)batch * n, 1.0, h->stream);
        if (verbose_timing) { B200_CUDA(cudaStreamSynchronize(h->stream)); }
        double t_1 = now();
        load_problem(h.get(), true, P, c, A, b, G, h_l, h_u, x_l, x_u, on_device);
        double t_2 = now();
        scale_problem(h.get(), false);
        if (verbose_timing) { B200_CUDA(cudaStreamSynchronize(h->stream)); }
        double t_3 = now();
        if (verbose_timing) fprintf(stderr, "[b200 setup] alloc %.1f ms, load(H2D+pack) %.1f ms, ruiz %.1f ms\n", t_1 - t_0, t_2 - t_1, t_3 - t_2);
        // hand the preconditioner to the IP driver (device pointers stay owned by RuizState)
        d.pd = h->ruiz.delta.get(); d.pd_inv = h->ruiz.delta_inv.get(); d.pdb = h->ruiz.delta_b.get(); d.pdb_inv = h->ruiz.delta_b_inv.get();
        d.pc = h->ruiz.c.get(); d.pc_inv = h->ruiz.c_inv.get();
        h->dense = std::make_unique<DenseBatchedKKT>(&h->dd, h->stream);
        h->be = h->dense.get();
        h->ip->finish_setup(h->be);
        B200_CUDA(cudaEventRecord(e1, h->stream));
        B200_CUDA(cudaEventSynchronize(e1));
        float ms = 0; B200_CUDA(cudaEventElapsedTime(&ms, e0, e1)); h->setup_ms = ms;
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        *out = h.release();
    )
    return B200_OK;
}

int b200qp_update_dense(b200qp_handle* h, const double* P, const double* c, const double* A, const double* b, const double* G,
                        const double* h_l, const double* h_u, const double* x_l, const double* x_u, int on_device) {
    if (!h) return fail(B200_E_INVALID, "null handle");
    if (h->kind != 0) return fail(B200_E_INVALID, "b200qp_update_dense on a sparse handle");
    B200_ZONE("piqp::Solver::update");
    B200_TRY(
        B200_CUDA(cudaSetDevice(h->device));
        IpDev& d = h->ip->dev();
        // solver.hpp:242: unscale, overwrite, rescale
        dense_ruiz_unscale(h->dd, h->ruiz, d.c, d.b, d.h_l, d.h_u, d.x_l, d.x_u, d.xbs, h->stream);
        int options = 0;
        if (P) options |= B200_KKT_UPDATE_P;
        if (A) options |= B200_KKT_UPDATE_A;
        if (G) options |= B200_KKT_UPDATE_G;
        load_problem(h, false, P, c, A, b, G, h_l, h_u, x_l, x_u, on_device);
        bool reuse = h->st.preconditioner_reuse_on_update != 0;
        if (options == 0) reuse = true;
        scale_problem(h, reuse);
        // a recomputed preconditioner rescales P, A and G alike: every cached product (AtA) is stale, not only those of the pieces
        // the caller passed (the reference refreshes only `options`, solver.hpp:290-301, and then factorises a KKT that
        // disagrees with its own data; deliberate deviation)
        if (!reuse) options = B200_KKT_UPDATE_P | B200_KKT_UPDATE_A | B200_KKT_UPDATE_G;
        else if (h_l || h_u) options |= B200_KKT_UPDATE_G;   // disable_inf_constraints may have zeroed rows of G (dense/data.hpp:144-169)
        h->dense->update_data(options);
        h->ip->finish_setup(h->dense.get());
        B200_CUDA(cudaStreamSynchronize(h->stream));
    )
    return B200_OK;
}

namespace {
__global__ void k_gather_vals(const double* src, long long src_stride, const int* map, int nnz, double* dst) {
    const int b = blockIdx.y;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < nnz) dst[(size_t)b * nnz + q] = src[(size_t)b * src_stride + map[q]];
}
// CSC of M (rows x cols) -> CSC of M^T with the map "entry of M^T -> entry of M"
void transpose_pattern(int rows, int cols, const int* cp, const int* ri, std::vector<int>& tp, std::vector<int>& ti, std::vector<int>& map) {
    const int nnz = cp ? cp[cols] : 0;
    tp.assign(rows + 1, 0); ti.assign(nnz, 0); map.assign(nnz, 0);
    for (int q = 0; q < nnz; q++) tp[ri[q] + 1]++;
    for (int r = 0; r < rows; r++) tp[r + 1] += tp[r];
    std::vector<int> w(tp.begin(), tp.end() - 1);
    if (cp) for (int j = 0; j < cols; j++) for (int q = cp[j]; q < cp[j + 1]; q++) { const int t = w[ri[q]]++; ti[t] = j; map[t] = q; }
}
void gather_values(b200qp_handle* h, const double* src, int nnz_in, const DevBuf<int>& map, int nnz, double* dst, int on_device) {
    if (!src || nnz == 0) return;
    Staged s;
    const double* dsrc = s.get(src, (size_t)h->batch * nnz_in, on_device, h->stream);
    dim3 g(ceil_div(nnz, 256), h->batch);
    B200_LAUNCH(k_gather_vals, g, 256, 0, h->stream, dsrc, (long long)nnz_in, map.get(), nnz, dst);
    B200_CUDA(cudaStreamSynchronize(h->stream));
}
void load_vectors_and_bounds(b200qp_handle* h, bool first, const double* c, const double* b, const double* h_l, const double* h_u,
                             const double* x_l, const double* x_u, int on_device) {
    const int B = h->batch, n = h->n, p = h->p, m = h->m;
    cudaStream_t st = h->stream;
    IpDev& d = h->ip->dev();
    copy_vec(d.c, c, (size_t)B * n, on_device, st);
    copy_vec(d.b, b, (size_t)B * p, on_device, st);
    Staged s1, s2, s3, s4;
    const double* dhl = s1.get(h_l, (size_t)B * m, on_device, st);
    const double* dhu = s2.get(h_u, (size_t)B * m, on_device, st);
    const double* dxl = s3.get(x_l, (size_t)B * n, on_device, st);
    const double* dxu = s4.get(x_u, (size_t)B * n, on_device, st);
    const int set_hl = first || h_l, set_hu = first || h_u, set_xl = first || x_l, set_xu = first || x_u;
    const int len = std::max(1, std::max(n, m));
    dim3 grid(ceil_div(len, 256), B);
    B200_LAUNCH(k_setup_bounds, grid, 256, 0, st, d, dhl, dhu, dxl, dxu, set_hl, set_hu, set_xl, set_xu, h->zero_rows.get());
    if (m > 0 && (set_hl || set_hu)) sparse_zero_G_rows(h->sd, h->zero_rows.get(), st);
    B200_CUDA(cudaStreamSynchronize(st));
}
}  // namespace

int b200qp_setup_sparse(b200qp_handle** out, int batch, int n, int p, int m,
                        const int* Pp, const int* Pi, const double* Px, const double* c,
                        const int* Ap, const int* Ai, const double* Ax, const double* b,
                        const int* Gp, const int* Gi, const double* Gx, const double* h_l, const double* h_u,
                        const double* x_l, const double* x_u, const b200qp_settings* settings, int device, int on_device) {
    return b200qp_setup_sparse_ex(out, batch, n, p, m, Pp, Pi, Px, c, Ap, Ai, Ax, b, Gp, Gi, Gx, h_l, h_u, x_l, x_u, settings, device, on_device, nullptr);
}
int b200qp_get_sparse_perm(b200qp_handle* h, int* perm, int cap) {
    if (!h || !h->ldlt) return fail(B200_E_INVALID, "not a sparse_ldlt handle");
    const std::vector<int>& pv = h->ldlt->S.perm;
    if (perm) for (int i = 0; i < (int)pv.size() && i < cap; i++) perm[i] = pv[i];
    return (int)pv.size();
}
int b200qp_setup_sparse_ex(b200qp_handle** out, int batch, int n, int p, int m,
                           const int* Pp, const int* Pi, const double* Px, const double* c,
                           const int* Ap, const int* Ai, const double* Ax, const double* b,
                           const int* Gp, const int* Gi, const double* Gx, const double* h_l, const double* h_u,
                           const double* x_l, const double* x_u, const b200qp_settings* settings, int device, int on_device, const int* kkt_perm) {
    if (!out || batch <= 0 || n <= 0 || p < 0 || m < 0 || !Pp || !c) return fail(B200_E_INVALID, "b200qp_setup_sparse: bad arguments");
    if ((p > 0 && (!Ap || !b)) || (m > 0 && (!Gp || (!h_l && !h_u)))) return fail(B200_E_INVALID, "b200qp_setup_sparse: missing constraint data");
    if (batch > B200_MAX_BATCH) return fail(B200_E_UNSUPPORTED, "b200qp_setup_sparse: batch > 65535 (the batch index is a gridDim.y / .z coordinate); split the batch over several handles");
    B200_ZONE("piqp::Solver::setup");
    B200_TRY(
        B200_CUDA(cudaSetDevice(device));
        const bool tm = getenv("B200_TIMING") != nullptr;
        auto now = [] { return std::chrono::steady_clock::now(); };
        auto t_0 = now(); auto t_prev = t_0;
        auto lap = [&](const char* what) { if (tm) { device_synchronize_shared(); auto t = now(); fprintf(stderr, "[b200qp_setup_sparse] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t - t_prev).count()); t_prev = t; } };
        auto h = std::make_unique<b200qp_handle>();
        h->kind = 1; h->device = device; h->batch = batch; h->n = n; h->p = p; h->m = m;
        if (settings) h->st = *settings; else b200qp_set_default_settings_sparse(&h->st);
        if (h->st.kkt_solver < 1 || h->st.kkt_solver > 5)
            throw std::runtime_error("b200qp_setup_sparse: kkt_solver must be one of the sparse backends (1..5, settings.hpp:18-26)");
        B200_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        cudaEvent_t e0, e1;
        B200_CUDA(cudaEventCreate(&e0)); B200_CUDA(cudaEventCreate(&e1));
        // patterns: P -> upper triangle (solver.hpp:182), A / G -> transposes (solver.hpp:183-184)
        std::vector<int> up, ui;
        up.assign(n + 1, 0);
        for (int j = 0; j < n; j++) { for (int q = Pp[j]; q < Pp[j + 1]; q++) if (Pi[q] <= j) { ui.push_back(Pi[q]); h->P_map.push_back(q); } up[j + 1] = (int)ui.size(); }
        h->nnzP_in = Pp[n]; h->nnzA_in = p > 0 ? Ap[n] : 0; h->nnzG_in = m > 0 ? Gp[n] : 0;
        SparseData& S = h->sd;
        S.n = n; S.p = p; S.m = m;
        S.P.build(n, n, up.data(), ui.data());
        std::vector<int> tp, ti;
        transpose_pattern(p, n, p > 0 ? Ap : nullptr, Ai, tp, ti, h->A_map);
        if (p == 0) tp.assign(1, 0);
        S.AT.build(n, p, tp.data(), ti.data());
        transpose_pattern(m, n, m > 0 ? Gp : nullptr, Gi, tp, ti, h->G_map);
        if (m == 0) tp.assign(1, 0);
        S.GT.build(n, m, tp.data(), ti.data());
        S.alloc_values(batch);
        auto upm = [](DevBuf<int>& d, const std::vector<int>& v) { d.alloc(std::max<size_t>(v.size(), 1)); if (!v.empty()) B200_CUDA(cudaMemcpy(d.get(), v.data(), v.size() * sizeof(int), cudaMemcpyHostToDevice)); };
        upm(h->d_P_map, h->P_map); upm(h->d_A_map, h->A_map); upm(h->d_G_map, h->G_map);
        B200_CUDA(device_synchronize_shared());
        lap("patterns + value buffers");
        B200_CUDA(cudaEventRecord(e0, h->stream));
        h->ip = std::make_unique<BatchedIPSolver>(batch, n, p, m, h->st, h->stream);
        h->ruiz.alloc(batch, n, p, m);
        lap("IP solver + ruiz alloc");
        h->zero_rows.alloc(std::max<size_t>((size_t)batch * m, 1)); h->zero_rows.zero(h->stream);
        IpDev& d = h->ip->dev();
        fill_d(d.xbs, (size_t)batch * n, 1.0, h->stream);
        gather_values(h.get(), Px, h->nnzP_in, h->d_P_map, S.P.nnz, S.Px.get(), on_device);
        gather_values(h.get(), Ax, h->nnzA_in, h->d_A_map, S.AT.nnz, S.ATx.get(), on_device);
        gather_values(h.get(), Gx, h->nnzG_in, h->d_G_map, S.GT.nnz, S.GTx.get(), on_device);
        load_vectors_and_bounds(h.get(), true, c, b, h_l, h_u, x_l, x_u, on_device);
        lap("H2D + gather");
        sparse_ruiz_scale(S, h->ruiz, d.c, d.b, d.h_l, d.h_u, d.x_l, d.x_u, d.xbs, false, h->st.preconditioner_scale_cost != 0, h->st.preconditioner_iter, h->stream);
        lap("ruiz");
        d.pd = h->ruiz.delta.get(); d.pd_inv = h->ruiz.delta_inv.get(); d.pdb = h->ruiz.delta_b.get(); d.pdb_inv = h->ruiz.delta_b_inv.get();
        d.pc = h->ruiz.c.get(); d.pc_inv = h->ruiz.c_inv.get();
        if (h->st.kkt_solver == 5) { h->ms = std::make_unique<MultistageBatchedKKT>(&h->sd, h->stream); h->be = h->ms.get(); }
        else { h->ldlt = std::make_unique<SparseLdltBatchedKKT>(&h->sd, kkt_perm, h->stream, h->st.kkt_solver - 1); h->be = h->ldlt.get(); }   // KKTMode = 0..3 (kkt_system.hpp:476-489)
        lap("backend ctor");
        h->ip->finish_setup(h->be);
        B200_CUDA(cudaEventRecord(e1, h->stream));
        B200_CUDA(cudaEventSynchronize(e1));
        lap("finish_setup");
        float ms_ = 0; B200_CUDA(cudaEventElapsedTime(&ms_, e0, e1)); h->setup_ms = ms_;
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        *out = h.release();
    )
    return B200_OK;
}

int b200qp_update_sparse(b200qp_handle* h, const double* Px, const double* c, const double* Ax, const double* b, const double* Gx,
                         const double* h_l, const double* h_u, const double* x_l, const double* x_u, int on_device) {
    if (!h) return fail(B200_E_INVALID, "null handle");
    if (h->kind != 1) return fail(B200_E_INVALID, "b200qp_update_sparse on a dense handle");
    B200_ZONE("piqp::Solver::update");
    B200_TRY(
        B200_CUDA(cudaSetDevice(h->device));
        IpDev& d = h->ip->dev();
        SparseData& S = h->sd;
        sparse_ruiz_unscale(S, h->ruiz, d.c, d.b, d.h_l, d.h_u, d.x_l, d.x_u, d.xbs, h->stream);
        int options = 0;
        if (Px) { options |= B200_KKT_UPDATE_P; gather_values(h, Px, h->nnzP_in, h->d_P_map, S.P.nnz, S.Px.get(), on_device); }
        if (Ax) { options |= B200_KKT_UPDATE_A; gather_values(h, Ax, h->nnzA_in, h->d_A_map, S.AT.nnz, S.ATx.get(), on_device); }
        if (Gx) { options |= B200_KKT_UPDATE_G; gather_values(h, Gx, h->nnzG_in, h->d_G_map, S.GT.nnz, S.GTx.get(), on_device); }
        load_vectors_and_bounds(h, false, c, b, h_l, h_u, x_l, x_u, on_device);
        bool reuse = h->st.preconditioner_reuse_on_update != 0;
        if (options == 0) reuse = true;
        sparse_ruiz_scale(S, h->ruiz, d.c, d.b, d.h_l, d.h_u, d.x_l, d.x_u, d.xbs, reuse, h->st.preconditioner_scale_cost != 0, h->st.preconditioner_iter, h->stream);
        if (!reuse) options = B200_KKT_UPDATE_P | B200_KKT_UPDATE_A | B200_KKT_UPDATE_G;   // see b200qp_update_dense
        else if (h_l || h_u) options |= B200_KKT_UPDATE_G;
        h->be->update_data(options);
        h->ip->finish_setup(h->be);
        B200_CUDA(cudaStreamSynchronize(h->stream));
    )
    return B200_OK;
}

int b200qp_multistage_blocks(b200qp_handle* h, int* out, int cap) {
    if (!h || !h->ms) return fail(B200_E_INVALID, "not a multistage handle");
    int k = 0;
    for (const auto& bl : h->ms->S.bi) { if (3 * k + 2 < cap) { out[3 * k] = bl.start; out[3 * k + 1] = bl.diag; out[3 * k + 2] = bl.off; } k++; }
    return k;
}
int b200kkt_multistage_blocks(b200kkt_handle* h, int* out, int cap) {
    if (!h || !h->ms) return fail(B200_E_INVALID, "not a multistage handle");
    int k = 0;
    for (const auto& bl : h->ms->S.bi) { if (3 * k + 2 < cap) { out[3 * k] = bl.start; out[3 * k + 1] = bl.diag; out[3 * k + 2] = bl.off; } k++; }
    return k;
}

int b200qp_update_settings(b200qp_handle* h, const b200qp_settings* s) {
    if (!h || !s) return fail(B200_E_INVALID, "null argument");
    h->st = *s; h->ip->set_settings(*s);
    return B200_OK;
}

int b200qp_solve(b200qp_handle* h) {
    if (!h) return fail(B200_E_INVALID, "null handle");
    B200_TRY(
        B200_CUDA(cudaSetDevice(h->device));
        h->be->reset_profile();
        h->ip->solve();
        h->solved = true;
    )
    return B200_OK;
}

int b200qp_get_result(b200qp_handle* h, double* x, double* y, double* z_l, double* z_u, double* z_bl, double* z_bu,
                      double* s_l, double* s_u, double* s_bl, double* s_bu, int on_device) {
    if (!h) return fail(B200_E_INVALID, "null handle");
    B200_TRY(
        B200_CUDA(cudaSetDevice(h->device));
        const Vars& v = h->ip->dev().it;
        const size_t B = h->batch;
        auto out = [&](double* dst, const double* src, size_t cnt) {
            if (dst && cnt) B200_CUDA(cudaMemcpyAsync(dst, src, cnt * sizeof(double), on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, h->stream));
        };
        out(x, v.x, B * h->n); out(y, v.y, B * h->p); out(z_l, v.z_l, B * h->m); out(z_u, v.z_u, B * h->m);
        out(z_bl, v.z_bl, B * h->n); out(z_bu, v.z_bu, B * h->n); out(s_l, v.s_l, B * h->m); out(s_u, v.s_u, B * h->m);
        out(s_bl, v.s_bl, B * h->n); out(s_bu, v.s_bu, B * h->n);
        B200_CUDA(cudaStreamSynchronize(h->stream));
    )
    return B200_OK;
}

int b200qp_get_info(b200qp_handle* h, b200qp_info* infos) {
    if (!h || !infos) return fail(B200_E_INVALID, "null argument");
    B200_TRY(
        B200_CUDA(cudaSetDevice(h->device));
        auto v = h->ip->infos();
        for (int b = 0; b < h->batch; b++) { infos[b] = v[b]; infos[b].setup_time = h->setup_ms * 1e-3; infos[b].run_time = infos[b].setup_time + infos[b].solve_time; }
    )
    return B200_OK;
}
int b200qp_get_stats(b200qp_handle* h, b200qp_stats* stats) {
    if (!h || !stats) return fail(B200_E_INVALID, "null argument");
    B200_TRY(
        h->ip->infos(); *stats = h->ip->stats();
        BatchedKKT* be = h->be;
        be->collect();
        stats->assemble_ms = be->prof_ms[BatchedKKT::T_ASSEMBLE]; stats->assemble_launches = be->prof_calls[BatchedKKT::T_ASSEMBLE];
        stats->cholesky_ms = be->prof_ms[BatchedKKT::T_FACTOR]; stats->cholesky_calls = be->prof_calls[BatchedKKT::T_FACTOR];
        stats->backend_solve_ms = be->prof_ms[BatchedKKT::T_SOLVE]; stats->backend_solve_launch_groups = be->prof_calls[BatchedKKT::T_SOLVE];
    )
    return B200_OK;
}
int b200qp_get_work(b200qp_handle* h, double* factor_flops, double* factor_bytes, double* solve_flops, double* solve_bytes) {
    if (!h || !h->be) return fail(B200_E_INVALID, "null handle");
    if (factor_flops) *factor_flops = h->be->factor_flops();
    if (factor_bytes) *factor_bytes = h->be->factor_bytes();
    if (solve_flops) *solve_flops = h->be->solve_flops();
    if (solve_bytes) *solve_bytes = h->be->solve_bytes();
    return B200_OK;
}
int b200qp_set_profiling(b200qp_handle* h, int enable) {
    if (!h) return fail(B200_E_INVALID, "null handle");
    h->be->collect();
    h->be->profile = enable != 0;
    return B200_OK;
}
int b200qp_get_trace(b200qp_handle* h, int b, double* rows, int max_rows) {
    if (!h || b < 0 || b >= h->batch) return fail(B200_E_INVALID, "bad instance");
    IpDev& d = h->ip->dev();
    if (!d.trace) return 0;
    int nrows = 0;
    B200_TRY(
        auto v = h->ip->infos();
        nrows = std::min(max_rows, std::min(d.trace_rows, v[b].iter + 1));
        B200_CUDA(cudaMemcpy(rows, d.trace + (size_t)b * d.trace_rows * 10, sizeof(double) * 10 * nrows, cudaMemcpyDeviceToHost));
    )
    return nrows;
}
void b200qp_cleanup(b200qp_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    // every API call returns with the handle's stream drained (its side streams join it before that), so nothing that touches the
    // handle's buffers is in flight after this one synchronisation
    if (h->stream) cudaStreamSynchronize(h->stream);
    ReleaseScope scope;
    delete h;
}

int b200qp_bench_factor_solve(b200qp_handle* h, int reps, int nsolve, double* factor_ms, double* solve_ms) {
    if (!h || reps <= 0) return fail(B200_E_INVALID, "bad arguments");
    B200_TRY(
        B200_CUDA(cudaSetDevice(h->device));
        IpDev& d = h->ip->dev();
        cudaEvent_t e[3];
        for (auto& x : e) B200_CUDA(cudaEventCreate(&x));
        double tf = 0, ts = 0;
        for (int r = 0; r < reps; r++) {
            B200_CUDA(cudaEventRecord(e[0], h->stream));
            h->be->factor(d.delta_reg, d.x_reg, d.z_reg_ir, nullptr, d.ok);
            B200_CUDA(cudaEventRecord(e[1], h->stream));
            for (int s = 0; s < nsolve; s++) h->be->solve(d.rhs_x_bar, d.r.y, d.rhs_z_bar, d.ref_x, d.ref_y, d.ref_z, nullptr);
            B200_CUDA(cudaEventRecord(e[2], h->stream));
            B200_CUDA(cudaEventSynchronize(e[2]));
            float a, b2;
            B200_CUDA(cudaEventElapsedTime(&a, e[0], e[1])); B200_CUDA(cudaEventElapsedTime(&b2, e[1], e[2]));
            tf += a; ts += b2;
        }
        for (auto& x : e) cudaEventDestroy(x);
        if (factor_ms) *factor_ms = tf / reps;
        if (solve_ms) *solve_ms = nsolve ? ts / (reps * nsolve) : 0.0;
    )
    return B200_OK;
}

}  // extern "C"
