// include/piqp_b200_adapter.hpp -- the reference-side adapter: libpiqp_b200 behind piqp::KKTSolverBase.
//
// Drop this header next to the reference's headers (it includes "piqp/kkt_solver_base.hpp", "piqp/dense/data.hpp",
// "piqp/sparse/data.hpp") and link -lpiqp_b200.  Every class below implements the 7 virtuals of
//   template<typename T, typename I, int MatrixType> class KKTSolverBase      (include/piqp/kkt_solver_base.hpp:21-44)
// by forwarding to the C-ABI of include/piqp_b200.h; nothing else of the reference changes except one `case` per backend in
// KKTSystem::init_kkt_solver (include/piqp/kkt_system.hpp:455-497) and the enum values of Settings::kkt_solver
// (include/piqp/settings.hpp:18-40) -- see INTEGRATION.md section 2.
//
// The reference cannot be compiled in this repository's build image (Eigen is absent), so tests/test_adapter_header.py compiles
// this header against a minimal mock of the four reference types it touches (tests/adapter_mock/piqp/*.hpp: KKTSolverBase, Vec,
// dense::Data, sparse::Data with Eigen's data() / outerIndexPtr() / innerIndexPtr() / valuePtr() accessors) and runs the dense
// adapter on the GPU box through it.
#ifndef PIQP_B200_ADAPTER_HPP
#define PIQP_B200_ADAPTER_HPP

#include <memory>

#include "piqp/kkt_solver_base.hpp"
#include "piqp/dense/data.hpp"
#include "piqp/sparse/data.hpp"
#include "piqp_b200.h"

namespace piqp {
namespace b200 {

namespace detail {
// the seven forwarding calls are identical for the three backends: only construction differs
template<typename T, typename I, int MatrixType, typename DataT>
class KKTAdapterBase : public KKTSolverBase<T, I, MatrixType> {
    static_assert(sizeof(T) == sizeof(double), "libpiqp_b200 computes in fp64 (include/piqp/common.hpp:38-39)");
protected:
    b200kkt_handle* h = nullptr;
    KKTAdapterBase() = default;
    KKTAdapterBase(const KKTAdapterBase& o) : h(o.h ? b200kkt_clone(o.h) : nullptr) {}
public:
    ~KKTAdapterBase() override { if (h) b200kkt_destroy(h); }
    bool ok() const { return h != nullptr; }      // construction failure: KKTSystem::init reports it (kkt_system.hpp:464,493)

    bool update_scalings_and_factor(const DataT&, const T& delta, const Vec<T>& x_reg, const Vec<T>& z_reg) override {
        return b200kkt_factor(h, delta, x_reg.data(), z_reg.data()) == 1;      // false = factorisation failed, never throws
    }
    void solve(const DataT&, const Vec<T>& rhs_x, const Vec<T>& rhs_y, const Vec<T>& rhs_z, Vec<T>& lhs_x, Vec<T>& lhs_y, Vec<T>& lhs_z) override {
        // pointers are re-read on every call: KKTSystem swaps its lhs buffers by pointer (kkt_system.hpp:292-300)
        b200kkt_solve(h, rhs_x.data(), rhs_y.data(), rhs_z.data(), lhs_x.data(), lhs_y.data(), lhs_z.data());
    }
    void eval_P_x(const DataT&, const T& alpha, const Vec<T>& x, Vec<T>& z) override { b200kkt_eval_P_x(h, alpha, x.data(), z.data()); }
    void eval_A_xn_and_AT_xt(const DataT&, const T& alpha_n, const T& alpha_t, const Vec<T>& xn, const Vec<T>& xt, Vec<T>& zn, Vec<T>& zt) override {
        b200kkt_eval_A_xn_and_AT_xt(h, alpha_n, alpha_t, xn.data(), xt.data(), zn.data(), zt.data());
    }
    void eval_G_xn_and_GT_xt(const DataT&, const T& alpha_n, const T& alpha_t, const Vec<T>& xn, const Vec<T>& xt, Vec<T>& zn, Vec<T>& zt) override {
        b200kkt_eval_G_xn_and_GT_xt(h, alpha_n, alpha_t, xn.data(), xt.data(), zn.data(), zt.data());
    }
    void print_info() override { b200kkt_print_info(h); }
};
}  // namespace detail

// ---- dense_cholesky  (replaces dense::KKT<T>, include/piqp/dense/kkt.hpp:26-178)
template<typename T>
class DenseKKT : public detail::KKTAdapterBase<T, int, PIQP_DENSE, dense::Data<T>> {
    using Base = detail::KKTAdapterBase<T, int, PIQP_DENSE, dense::Data<T>>;
public:
    explicit DenseKKT(const dense::Data<T>& d, int device = 0) {
        // dense::Data stores P_utri (n x n), AT (n x p), GT (n x m) column-major (dense/data.hpp:29-32): what the ABI takes
        if (b200kkt_dense_create(&this->h, int(d.n), int(d.p), int(d.m), d.P_utri.data(), d.AT.data(), d.GT.data(), device) != B200_OK) {
            piqp_eprint("b200: %s\n", b200_last_error());
            this->h = nullptr;
        }
    }
    DenseKKT(const DenseKKT& o) : Base(o) {}
    std::unique_ptr<KKTSolverBase<T, int, PIQP_DENSE>> clone() const override { return std::make_unique<DenseKKT>(*this); }
    void update_data(const dense::Data<T>& d, int options) override {
        b200kkt_update_data(this->h, options, d.P_utri.data(), d.AT.data(), d.GT.data());
    }
};

// ---- sparse_ldlt and its condensed variants  (replaces sparse::KKT<T, I, Mode>, include/piqp/sparse/kkt.hpp:31-250)
// Mode = KKT_FULL / KKT_EQ_ELIMINATED / KKT_INEQ_ELIMINATED / KKT_ALL_ELIMINATED (include/piqp/kkt_fwd.hpp:15-21): same integers as the ABI
template<typename T, typename I, int Mode>
class SparseKKT : public detail::KKTAdapterBase<T, I, PIQP_SPARSE, sparse::Data<T, I>> {
    static_assert(sizeof(I) == sizeof(int), "libpiqp_b200 uses int32 indices (include/piqp/common.hpp:38-39)");
    using Base = detail::KKTAdapterBase<T, I, PIQP_SPARSE, sparse::Data<T, I>>;
public:
    explicit SparseKKT(const sparse::Data<T, I>& d, int device = 0) {
        if (b200kkt_sparse_create(&this->h, int(d.n), int(d.p), int(d.m),
                                  d.P_utri.outerIndexPtr(), d.P_utri.innerIndexPtr(), d.P_utri.valuePtr(),
                                  d.AT.outerIndexPtr(), d.AT.innerIndexPtr(), d.AT.valuePtr(),
                                  d.GT.outerIndexPtr(), d.GT.innerIndexPtr(), d.GT.valuePtr(), Mode, /*perm=*/nullptr, device) != B200_OK) {
            piqp_eprint("b200: %s\n", b200_last_error());
            this->h = nullptr;
        }
    }
    SparseKKT(const SparseKKT& o) : Base(o) {}
    std::unique_ptr<KKTSolverBase<T, I, PIQP_SPARSE>> clone() const override { return std::make_unique<SparseKKT>(*this); }
    void update_data(const sparse::Data<T, I>& d, int options) override {      // same pattern, new values (sparse/kkt_full.hpp:212-251)
        b200kkt_update_data(this->h, options, d.P_utri.valuePtr(), d.AT.valuePtr(), d.GT.valuePtr());
    }
};

// ---- sparse_multistage  (replaces sparse::MultistageKKT<T, I>, include/piqp/sparse/multistage_kkt.hpp:41-1816; no BLASFEO needed)
template<typename T, typename I>
class MultistageKKT : public detail::KKTAdapterBase<T, I, PIQP_SPARSE, sparse::Data<T, I>> {
    static_assert(sizeof(I) == sizeof(int), "libpiqp_b200 uses int32 indices");
    using Base = detail::KKTAdapterBase<T, I, PIQP_SPARSE, sparse::Data<T, I>>;
public:
    explicit MultistageKKT(const sparse::Data<T, I>& d, int device = 0) {
        if (b200kkt_multistage_create(&this->h, int(d.n), int(d.p), int(d.m),
                                      d.P_utri.outerIndexPtr(), d.P_utri.innerIndexPtr(), d.P_utri.valuePtr(),
                                      d.AT.outerIndexPtr(), d.AT.innerIndexPtr(), d.AT.valuePtr(),
                                      d.GT.outerIndexPtr(), d.GT.innerIndexPtr(), d.GT.valuePtr(), device) != B200_OK) {
            piqp_eprint("b200: %s\n", b200_last_error());
            this->h = nullptr;
        }
    }
    MultistageKKT(const MultistageKKT& o) : Base(o) {}
    std::unique_ptr<KKTSolverBase<T, I, PIQP_SPARSE>> clone() const override { return std::make_unique<MultistageKKT>(*this); }
    void update_data(const sparse::Data<T, I>& d, int options) override {
        b200kkt_update_data(this->h, options, d.P_utri.valuePtr(), d.AT.valuePtr(), d.GT.valuePtr());
    }
};

}  // namespace b200
}  // namespace piqp

#endif  // PIQP_B200_ADAPTER_HPP
