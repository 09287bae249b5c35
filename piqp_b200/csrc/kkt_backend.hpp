// piqp_b200/csrc/kkt_backend.hpp -- device-side twin of piqp::KKTSolverBase for a BATCH of instances.
//
// The reference's plugin interface (include/piqp/kkt_solver_base.hpp:21-44) has seven virtuals that act on
// one QP.  Here the same seven operations act on `batch` independent QPs of identical shape whose vectors
// live in HBM, instance-major ([batch][len], no padding).  Every call takes an optional per-instance
// `active` mask so the lock-step interior-point driver can retire converged instances.
#pragma once
#include <vector>
#include "common.cuh"

namespace b200 {

struct BatchedKKT {
    int batch = 0, n = 0, p = 0, m = 0;
    cudaStream_t stream = 0;
    virtual ~BatchedKKT() = default;

    // KKTSolverBase::update_data(data, options): the owner has already rewritten the device problem data
    virtual void update_data(int options) = 0;
    // KKTSolverBase::update_scalings_and_factor: delta[batch], x_reg[batch][n], z_reg[batch][m];
    // ok[b] <- 1 / 0 for active instances
    virtual void factor(const double* delta, const double* x_reg, const double* z_reg, const int* active, int* ok) = 0;
    // KKTSolverBase::solve
    virtual void solve(const double* rx, const double* ry, const double* rz, double* lx, double* ly, double* lz, const int* active) = 0;
    // z = alpha P x
    virtual void eval_P_x(double alpha, const double* x, double* z, const int* active) = 0;
    // zn = an * A xn ; zt = at * A^T xt
    virtual void eval_A(double an, double at, const double* xn, const double* xt, double* zn, double* zt, const int* active) = 0;
    virtual void eval_G(double an, double at, const double* xn, const double* xt, double* zn, double* zt, const int* active) = 0;
    // diag(P) for KKTSystem's static regularisation (kkt_system.hpp:198,430-453): P_diag[batch][n]
    virtual void extract_P_diag(double* P_diag) = 0;
    virtual void print_info() const {}
    // true if factor / solve / eval_* only enqueue work on `stream` (no host synchronisation, no other streams): the IP driver
    // may then capture a whole iteration into a CUDA graph
    virtual bool graph_capturable() const { return false; }
    // true if factor() reports ok = 1 for every active instance, whatever the data (the IP driver may then enqueue the next iteration
    // before it has read the flags of the current one)
    virtual bool factor_never_fails() const { return false; }
    // algorithmic work per call and instance, for GFLOP/s and roofline reporting
    virtual double factor_flops() const = 0;
    virtual double factor_bytes() const = 0;
    virtual double solve_flops() const = 0;
    virtual double solve_bytes() const = 0;

    // ---- optional per-kernel-class device timers (CUDA events on `stream`), used by bench.py's roofline block
    enum { T_ASSEMBLE = 0, T_FACTOR = 1, T_SOLVE = 2, T_COUNT = 3 };
    bool profile = false;
    double prof_ms[T_COUNT] = {0, 0, 0};
    long long prof_calls[T_COUNT] = {0, 0, 0};
    void tic(int kind);
    void toc(int kind);
    void collect();      // after a stream sync: fold finished event pairs into prof_ms / prof_calls
    void reset_profile();
private:
    struct Span { cudaEvent_t a, b; int kind; };
    std::vector<Span> spans_;
    std::vector<cudaEvent_t> free_events_;
    cudaEvent_t open_[T_COUNT] = {nullptr, nullptr, nullptr};
};

}  // namespace b200
