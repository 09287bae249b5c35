// piqp_b200/csrc/ip_solver.hpp -- batched, device-resident interior-point driver.
//
// Restates, per instance and on the GPU, what the reference runs on the host around its KKT backend:
//   KKTSystem::update_scalings_and_factor / solve (+ iterative refinement)   include/piqp/kkt_system.hpp:143-369,499-536
//   SolverBase::solve_impl and its residual / step helpers                   include/piqp/solver.hpp:379-1259
// All instances of a batch advance in lock step; per-instance control flow (factor retry ladder,
// refinement trip counts, termination) is expressed with device-side masks and one small flag
// read-back per phase.  The backend behind it is any BatchedKKT (dense / sparse / multistage).
#pragma once
#include <memory>
#include <vector>
#include "../../include/piqp_b200.h"
#include "kkt_backend.hpp"

namespace b200 {

struct Vars {   // device twin of piqp::Variables (include/piqp/variables.hpp:63-105); box blocks are x-indexed
    double *x, *y, *z_l, *z_u, *z_bl, *z_bu, *s_l, *s_u, *s_bl, *s_bu;
};

struct IpScalars {   // per-instance scalar state == piqp::Info (include/piqp/results.hpp:45-89) + driver flags
    int status, iter, factor_retires, no_primal_update, no_dual_update;
    int active, ir_on, need_factor, reg_changed, has_ineq, use_ir, ir_continue;
    double rho, delta, mu, sigma, primal_step, dual_step;
    double primal_res, primal_res_rel, dual_res, dual_res_rel;
    double primal_res_reg, primal_res_reg_rel, dual_res_reg, dual_res_reg_rel;
    double primal_prox_inf, dual_prox_inf, prev_primal_res, prev_dual_res;
    double primal_obj, dual_obj, duality_gap, duality_gap_rel, reg_limit;
    double kkt_rho, kkt_delta, n_fin, mu_rate;
    double rhs_norm, refine_err;
    long long n_factor, n_solve, n_backend_solve;
};

struct IpDev {   // everything a phase kernel needs, passed by value
    int batch, n, p, m;
    // problem vectors (scaled in place by the preconditioner); box vectors are x-indexed
    double *c, *b, *h_l, *h_u, *x_l, *x_u, *xbs;
    int *has_hl, *has_hu, *has_xl, *has_xu;
    // preconditioner (all ones for the identity preconditioner)
    double *pd, *pd_inv, *pdb, *pdb_inv, *pc, *pc_inv;
    Vars it, r, rnr, step, prox;
    // KKTSystem members (kkt_system.hpp:32-62)
    double *k_s_l, *k_s_u, *k_s_bl, *k_s_bu, *k_zl_inv, *k_zu_inv, *k_zbl_inv, *k_zbu_inv;
    double *x_reg, *z_reg, *z_reg_ir, *rhs_x_bar, *rhs_z_bar, *lhs_z;
    double *err_x, *err_y, *err_z, *ref_x, *ref_y, *ref_z, *work_x, *work_x2, *work_z, *P_diag;
    IpScalars* sc;
    double* delta_reg;   // [batch] what the backend gets as delta
    int *act, *act2, *need_factor, *ok, *ir_mask;
    double* trace; int trace_rows;   // optional [batch][trace_rows][10]
    b200qp_settings st;
};

class BatchedIPSolver {
public:
    BatchedIPSolver(int batch, int n, int p, int m, const b200qp_settings& st, cudaStream_t stream);
    ~BatchedIPSolver();
    // The owner fills the problem vectors / masks (device pointers in dev()) and attaches a backend, then:
    void finish_setup(BatchedKKT* backend);   // counts, P_diag
    void solve();                             // SolverBase::solve(): solve_impl + unscale_results + restore_dual
    IpDev& dev() { return d_; }
    std::vector<b200qp_info> infos();
    b200qp_stats stats() const { return stats_; }
    void set_settings(const b200qp_settings& st) { d_.st = st; drop_graph(); }      // settings travel in the kernel arguments: re-capture
    void drop_graph();
    int batch, n, p, m;
    cudaStream_t stream;
    bool identity_precond = false;

private:
    void kkt_solve(const Vars& rhs, const Vars& lhs, const int* mask);   // KKTSystem::solve
    int factor_with_retry(bool first_round_done);      // returns the number of instances still active (read back with the factor flags: one host sync)
    void factor_round();
    int read_factor_flags(int& active);
    void iteration_body(cudaEvent_t after_solves);
    bool ensure_graph();
    cudaGraphExec_t graph_exec_ = nullptr;
    unsigned long long graph_launches_ = 0;
    bool graph_failed_ = false, use_graphs_ = true, bucket_timers_ = true;
    void residuals_nr(const int* mask);
    int count_flags(const int* dev_flags, int count = -1);
    IpDev d_{};
    BatchedKKT* be_ = nullptr;
    std::vector<DevBuf<double>> pool_;
    std::vector<DevBuf<int>> ipool_;
    double* arena_ptr_ = nullptr;      // bump allocator over the last slab of pool_ (alloc_d / alloc_i)
    size_t arena_left_ = 0;
    DevBuf<IpScalars> sc_;
    int* h_flags_ = nullptr;   // pinned
    int* h_flags2_[2] = {nullptr, nullptr};      // pinned, double-buffered read-back of pipelined graph replays
    cudaEvent_t flag_ev_[2] = {nullptr, nullptr};
    b200qp_stats stats_{};
    cudaEvent_t ev_[6];
    int ipt_ = 256;
    std::vector<cudaEvent_t> iter_ev_;   // 3 per IP iteration (factor begin, factor end / solve begin, solve end): read after the loop, no per-iteration sync
    bool any_ir_ = false, invalid_settings_ = false;
    int* ir_was_ = nullptr;
    double* alloc_d(size_t n);
    int* alloc_i(size_t n);
    Vars alloc_vars();
};

}  // namespace b200
