// oracle/oracle_core.hpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement (plain C++17, no Eigen) of the parts of PIQP v0.6.2 that sit
// AROUND the KKT hot path: problem data bookkeeping, Ruiz equilibration, the
// KKTSystem reduction layer (slack/box elimination, iterative refinement, dual
// recovery) and the interior-point loop.  It exists so that tests can
//   (1) pin the algorithm against the reference's known-answer QPs, and
//   (2) drive either the oracle's CPU backends or the CUDA backend (through the
//       C-ABI of libpiqp_b200) with the SAME caller, like the reference solver would.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use it.
//
// Parity status: "parity unpinned" at the factor-entry level (the reference cannot be
// compiled here: Eigen is absent); pinned at solver level by the reference's
// known-answer tests (tests/src/dense/solver_test.cpp, sparse/solver_test.cpp) and the
// notebook trace (docs/assets/robust_scenario_mpc.ipynb:489-573).
//
// Reference files followed (relative to /root/reference/include/piqp):
//   settings.hpp:43-107, results.hpp:18-95, variables.hpp:17-105,
//   dense/data.hpp:100-212, dense/preconditioner.hpp:45-437,
//   kkt_system.hpp:97-369,499-536, solver.hpp:151-216,379-1259.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <memory>
#include <vector>

namespace oracle {

using Vec = std::vector<double>;
using IVec = std::vector<int>;
constexpr double kInf = 1e30;  // fwd.hpp:54 (PIQP_INF)

// settings.hpp:43-82
struct Settings {
    double rho_init = 1e-6;
    double delta_init = 1e-4;
    double eps_abs = 1e-8;
    double eps_rel = 1e-9;
    int check_duality_gap = 1;
    double eps_duality_gap_abs = 1e-8;
    double eps_duality_gap_rel = 1e-9;
    double infeasibility_threshold = 0.9;
    double reg_lower_limit = 1e-10;
    double reg_finetune_lower_limit = 1e-13;
    long reg_finetune_primal_update_threshold = 7;
    long reg_finetune_dual_update_threshold = 7;
    long max_iter = 250;
    long max_factor_retires = 10;
    int preconditioner_scale_cost = 0;
    int preconditioner_reuse_on_update = 0;
    long preconditioner_iter = 10;
    double tau = 0.99;
    int kkt_solver = 0;
    int iterative_refinement_always_enabled = 0;
    double iterative_refinement_eps_abs = 1e-12;
    double iterative_refinement_eps_rel = 1e-12;
    long iterative_refinement_max_iter = 10;
    double iterative_refinement_min_improvement_rate = 5.0;
    double iterative_refinement_static_regularization_eps = 1e-8;
    double iterative_refinement_static_regularization_rel =
        std::numeric_limits<double>::epsilon() * std::numeric_limits<double>::epsilon();
    int verbose = 0;
    int compute_timings = 0;

    // settings.hpp:84-106
    bool verify() const {
        return rho_init > 0 && delta_init > 0 && eps_abs > 0 && eps_rel >= 0 &&
               eps_duality_gap_abs > 0 && eps_duality_gap_rel >= 0 &&
               infeasibility_threshold >= 0 && reg_lower_limit > 0 &&
               reg_finetune_primal_update_threshold >= 0 &&
               reg_finetune_dual_update_threshold >= 0 && max_iter > 0 &&
               max_factor_retires > 0 && preconditioner_iter >= 0 && tau > 0 && tau <= 1 &&
               iterative_refinement_eps_abs > 0 && iterative_refinement_eps_rel >= 0 &&
               iterative_refinement_max_iter >= 0 &&
               iterative_refinement_min_improvement_rate >= 1.0 &&
               iterative_refinement_static_regularization_eps > 0 &&
               iterative_refinement_static_regularization_rel >= 0;
    }
};

// results.hpp:18-27
enum Status {
    SOLVED = 1,
    MAX_ITER_REACHED = -1,
    PRIMAL_INFEASIBLE = -2,
    DUAL_INFEASIBLE = -3,
    NUMERICS = -8,
    UNSOLVED = -9,
    INVALID_SETTINGS = -10
};

// results.hpp:45-89
struct Info {
    int status = UNSOLVED;
    long iter = 0;
    double rho = 0, delta = 0, mu = 0, sigma = 0, primal_step = 0, dual_step = 0;
    double primal_res = 0, primal_res_rel = 0, dual_res = 0, dual_res_rel = 0;
    double primal_res_reg = 0, primal_res_reg_rel = 0, dual_res_reg = 0, dual_res_reg_rel = 0;
    double primal_prox_inf = 0, dual_prox_inf = 0;
    double prev_primal_res = 0, prev_dual_res = 0;
    double primal_obj = 0, dual_obj = 0, duality_gap = 0, duality_gap_rel = 0;
    long factor_retires = 0;
    double reg_limit = 0;
    long no_primal_update = 0, no_dual_update = 0;
    double setup_time = 0, update_time = 0, solve_time = 0, kkt_factor_time = 0,
           kkt_solve_time = 0, run_time = 0;
    // extra counters (ours, for throughput accounting; not in the reference)
    long n_factor = 0, n_solve = 0, n_backend_solve = 0;
};

// variables.hpp:17-105
struct Variables {
    Vec x, y, z_l, z_u, z_bl, z_bu, s_l, s_u, s_bl, s_bu;
    void resize(int n, int p, int m) {
        x.assign(n, 0); y.assign(p, 0);
        z_l.assign(m, 0); z_u.assign(m, 0); z_bl.assign(n, 0); z_bu.assign(n, 0);
        s_l.assign(m, 0); s_u.assign(m, 0); s_bl.assign(n, 0); s_bu.assign(n, 0);
    }
};

// kkt_fwd.hpp:23-29
enum UpdateOptions { UPDATE_NONE = 0, UPDATE_P = 1, UPDATE_A = 2, UPDATE_G = 4 };

// The vector part of dense::Data / sparse::Data (dense/data.hpp:23-51, identical fields in
// sparse/data.hpp:26-54).  Matrices live behind QPMatrices.
struct ProblemVectors {
    int n = 0, p = 0, m = 0;
    Vec c, b, h_l, h_u, x_l, x_u;
    int n_h_l = 0, n_h_u = 0, n_x_l = 0, n_x_u = 0;
    IVec h_l_idx, h_u_idx, x_l_idx, x_u_idx;
    Vec x_b_scaling;

    void resize(int n_, int p_, int m_) {  // dense/data.hpp:74-98
        n = n_; p = p_; m = m_;
        c.assign(n, 0); b.assign(p, 0); h_l.assign(m, 0); h_u.assign(m, 0);
        x_l.assign(n, 0); x_u.assign(n, 0);
        h_l_idx.assign(m, 0); h_u_idx.assign(m, 0); x_l_idx.assign(n, 0); x_u_idx.assign(n, 0);
        x_b_scaling.assign(n, 1.0);
    }
    // dense/data.hpp:100-119 ; null pointer == nullopt
    void set_h_l(const double* v) {
        n_h_l = 0;
        if (v) {
            int k = 0;
            for (int i = 0; i < m; i++) {
                if (v[i] > -kInf) { n_h_l++; h_l[i] = v[i]; h_l_idx[k++] = i; }
                else h_l[i] = -kInf;
            }
        } else std::fill(h_l.begin(), h_l.end(), -kInf);
    }
    // dense/data.hpp:121-140
    void set_h_u(const double* v) {
        n_h_u = 0;
        if (v) {
            int k = 0;
            for (int i = 0; i < m; i++) {
                if (v[i] < kInf) { n_h_u++; h_u[i] = v[i]; h_u_idx[k++] = i; }
                else h_u[i] = kInf;
            }
        } else std::fill(h_u.begin(), h_u.end(), kInf);
    }
    // dense/data.hpp:171-186 (compact storage: finite bounds packed in the head)
    void set_x_l(const double* v) {
        n_x_l = 0;
        if (v) for (int i = 0; i < n; i++) if (v[i] > -kInf) { x_l[n_x_l] = v[i]; x_l_idx[n_x_l] = i; n_x_l++; }
    }
    void set_x_u(const double* v) {
        n_x_u = 0;
        if (v) for (int i = 0; i < n; i++) if (v[i] < kInf) { x_u[n_x_u] = v[i]; x_u_idx[n_x_u] = i; n_x_u++; }
    }
};

// The 7-method plugin interface, kkt_solver_base.hpp:21-44, on raw pointers.
struct KKTBackend {
    virtual ~KKTBackend() = default;
    virtual void update_data(int options) = 0;
    virtual bool factor(double delta, const double* x_reg, const double* z_reg) = 0;
    virtual void solve(const double* rx, const double* ry, const double* rz, double* lx, double* ly, double* lz) = 0;
    virtual void eval_P_x(double alpha, const double* x, double* z) = 0;
    virtual void eval_A(double an, double at, const double* xn, const double* xt, double* zn, double* zt) = 0;
    virtual void eval_G(double an, double at, const double* xn, const double* xt, double* zn, double* zt) = 0;
};

// Matrix storage + the matrix-touching pieces of Ruiz equilibration.
struct QPMatrices {
    virtual ~QPMatrices() = default;
    virtual void set_G_row_zero(int row) = 0;                 // dense/data.hpp:208-211
    // fills nrm[0..n+p+m) with the inf-norms of the columns of [P AT GT; A 0 0; G 0 0]
    // (x_b_scaling is folded in by the caller for the dense case -- see Ruiz below)
    virtual void kkt_col_norms(const Vec& x_b_scaling, Vec& nrm) = 0;
    virtual void scale_sym(const double* d) = 0;              // P, AT, GT <- D P D, Dx AT Dy, Dx GT Dz
    virtual double cost_norm_mean() = 0;                      // gamma numerator of scale_cost
    virtual void scale_P(double gamma) = 0;
    virtual void extract_P_diag(Vec& d) = 0;                  // kkt_system.hpp:430-453
    virtual std::unique_ptr<KKTBackend> make_backend(int kkt_solver) = 0;
};

inline double inf_norm(const Vec& v) { double r = 0; for (double e : v) r = std::max(r, std::fabs(e)); return r; }
inline double inf_norm(const double* v, int n) { double r = 0; for (int i = 0; i < n; i++) r = std::max(r, std::fabs(v[i])); return r; }
inline double dot(const double* a, const double* b, int n) { double r = 0; for (int i = 0; i < n; i++) r += a[i] * b[i]; return r; }

// dense/preconditioner.hpp:26-437 (sparse/preconditioner.hpp:28-557 has the same arithmetic on CSC).
struct Ruiz {
    bool identity = false;   // IdentityPreconditioner, dense/preconditioner.hpp:440-560
    int n = 0, p = 0, m = 0;
    double c = 1, c_inv = 1;
    Vec delta, delta_b, delta_inv, delta_b_inv;

    static void limit(double& d) { if (d < 1e-4) d = 1.0; else if (d > 1e4) d = 1e4; }  // :424-436

    void init(const ProblemVectors& pv) {  // :43-61
        n = pv.n; p = pv.p; m = pv.m;
        delta.assign(n + p + m, 1.0); delta_b.assign(n, 1.0);
        delta_inv.assign(n + p + m, 1.0); delta_b_inv.assign(n, 1.0);
        c = 1; c_inv = 1;
    }

    void scale_data(ProblemVectors& pv, QPMatrices& M, bool reuse, bool scale_cost, long max_iter, double eps = 1e-3) {
        if (identity) return;
        const int N = n + p + m;
        if (!reuse) {  // :68-164
            c = 1; std::fill(delta.begin(), delta.end(), 1.0); std::fill(delta_b.begin(), delta_b.end(), 1.0);
            Vec& it = delta_inv; Vec& itb = delta_b_inv;  // temporaries, as in the reference
            std::fill(it.begin(), it.end(), 0.0); std::fill(itb.begin(), itb.end(), 0.0);
            for (long iter = 0; iter < max_iter; iter++) {
                double dev = 0;
                for (double e : it) dev = std::max(dev, std::fabs(1 - e));
                for (double e : itb) dev = std::max(dev, std::fabs(1 - e));
                if (!(dev > eps)) break;
                M.kkt_col_norms(pv.x_b_scaling, it);
                for (int k = 0; k < n; k++) itb[k] = pv.x_b_scaling[k];
                for (int k = 0; k < N; k++) limit(it[k]);
                for (int k = 0; k < n; k++) limit(itb[k]);
                for (int k = 0; k < N; k++) it[k] = 1.0 / std::sqrt(it[k]);
                for (int k = 0; k < n; k++) itb[k] = 1.0 / std::sqrt(itb[k]);
                M.scale_sym(it.data());
                for (int k = 0; k < n; k++) pv.c[k] *= it[k];
                for (int k = 0; k < n; k++) pv.x_b_scaling[k] *= itb[k] * it[k];
                for (int k = 0; k < N; k++) delta[k] *= it[k];
                for (int k = 0; k < n; k++) delta_b[k] *= itb[k];
                if (scale_cost) {  // :141-162
                    double gamma = M.cost_norm_mean();
                    limit(gamma);
                    gamma = std::max(gamma, inf_norm(pv.c));
                    limit(gamma);
                    gamma = 1.0 / gamma;
                    M.scale_P(gamma);
                    for (int k = 0; k < n; k++) pv.c[k] *= gamma;
                    c *= gamma;
                }
            }
            c_inv = 1.0 / c;
            for (int k = 0; k < N; k++) delta_inv[k] = 1.0 / delta[k];
            for (int k = 0; k < n; k++) delta_b_inv[k] = 1.0 / delta_b[k];
        } else {  // :166-197
            M.scale_P(c);
            M.scale_sym(delta.data());
            for (int k = 0; k < n; k++) pv.c[k] *= c * delta[k];
            for (int k = 0; k < n; k++) pv.x_b_scaling[k] *= delta_b[k] * delta[k];
        }
        // :199-214
        for (int k = 0; k < p; k++) pv.b[k] *= delta[n + k];
        for (int k = 0; k < m; k++) { pv.h_l[k] *= delta[n + p + k]; pv.h_u[k] *= delta[n + p + k]; }
        for (int i = 0; i < pv.n_x_l; i++) pv.x_l[i] *= delta_b[pv.x_l_idx[i]];
        for (int i = 0; i < pv.n_x_u; i++) pv.x_u[i] *= delta_b[pv.x_u_idx[i]];
    }

    void unscale_data(ProblemVectors& pv, QPMatrices& M) {  // :217-251
        if (identity) return;
        M.scale_P(c_inv);
        M.scale_sym(delta_inv.data());
        for (int k = 0; k < n; k++) pv.c[k] *= c_inv * delta_inv[k];
        for (int k = 0; k < n; k++) pv.x_b_scaling[k] *= delta_b_inv[k] * delta_inv[k];
        for (int k = 0; k < p; k++) pv.b[k] *= delta_inv[n + k];
        for (int k = 0; k < m; k++) { pv.h_l[k] *= delta_inv[n + p + k]; pv.h_u[k] *= delta_inv[n + p + k]; }
        for (int i = 0; i < pv.n_x_l; i++) pv.x_l[i] *= delta_b_inv[pv.x_l_idx[i]];
        for (int i = 0; i < pv.n_x_u; i++) pv.x_u[i] *= delta_b_inv[pv.x_u_idx[i]];
    }
    // element-wise (un)scalers used by the IP loop, :253-421
    double unscale_cost(double v) const { return c_inv * v; }
    double us_primal(double v, int i) const { return v * delta[i]; }
    double us_dual_eq(double v, int i) const { return v * c_inv * delta[n + i]; }
    double us_dual_ineq(double v, int i) const { return v * c_inv * delta[n + p + i]; }
    double us_dual_b(double v, int i) const { return v * c_inv * delta_b[i]; }
    double us_slack_ineq(double v, int i) const { return v * delta_inv[n + p + i]; }
    double us_slack_b(double v, int i) const { return v * delta_b_inv[i]; }
    double us_pres_eq(double v, int i) const { return v * delta_inv[n + i]; }
    double us_pres_ineq(double v, int i) const { return v * delta_inv[n + p + i]; }
    double us_pres_b(double v, int i) const { return v * delta_b_inv[i]; }
    double us_dres(double v, int i) const { return v * c_inv * delta_inv[i]; }
};

// kkt_system.hpp:27-537
struct KKTSystem {
    double rho = 0, delta = 0;
    Vec P_diag, s_l, s_u, s_bl, s_bu, z_l_inv, z_u_inv, z_bl_inv, z_bu_inv;
    Vec x_reg, z_reg, rhs_x_bar, rhs_z_bar, work_x, work_z;
    Vec err_x, err_y, err_z, ref_x, ref_y, ref_z;
    bool use_ir = false;
    std::unique_ptr<KKTBackend> be;
    long n_backend_solve = 0;

    bool init(const ProblemVectors& d, QPMatrices& M, const Settings& st) {  // :97-132
        P_diag.assign(d.n, 0);
        s_l.assign(d.m, 0); s_u.assign(d.m, 0); s_bl.assign(d.n, 0); s_bu.assign(d.n, 0);
        z_l_inv.assign(d.m, 0); z_u_inv.assign(d.m, 0); z_bl_inv.assign(d.n, 0); z_bu_inv.assign(d.n, 0);
        x_reg.assign(d.n, 0); z_reg.assign(d.m, 0); rhs_x_bar.assign(d.n, 0); rhs_z_bar.assign(d.m, 0);
        work_x.assign(d.n, 0); work_z.assign(d.m, 0);
        err_x.assign(d.n, 0); err_y.assign(d.p, 0); err_z.assign(d.m, 0);
        ref_x.assign(d.n, 0); ref_y.assign(d.p, 0); ref_z.assign(d.m, 0);
        M.extract_P_diag(P_diag);
        be = M.make_backend(st.kkt_solver);
        return be != nullptr;
    }
    void update_data(QPMatrices& M, int options) {  // :134-141
        if (options & UPDATE_P) M.extract_P_diag(P_diag);
        be->update_data(options);
    }

    // :143-211
    bool update_scalings_and_factor(const ProblemVectors& d, const Settings& st, bool ir, double rho_, double delta_, const Variables& v) {
        Vec& z_reg_ir = work_z;
        rho = rho_; delta = delta_;
        s_l = v.s_l; s_u = v.s_u;
        for (int i = 0; i < d.n_x_l; i++) s_bl[i] = v.s_bl[i];
        for (int i = 0; i < d.n_x_u; i++) s_bu[i] = v.s_bu[i];
        for (int i = 0; i < d.m; i++) { z_l_inv[i] = 1.0 / v.z_l[i]; z_u_inv[i] = 1.0 / v.z_u[i]; }
        for (int i = 0; i < d.n_x_l; i++) z_bl_inv[i] = 1.0 / v.z_bl[i];
        for (int i = 0; i < d.n_x_u; i++) z_bu_inv[i] = 1.0 / v.z_bu[i];

        std::fill(x_reg.begin(), x_reg.end(), rho);
        for (int i = 0; i < d.n_x_l; i++) { int k = d.x_l_idx[i]; x_reg[k] += d.x_b_scaling[k] * d.x_b_scaling[k] / (z_bl_inv[i] * s_bl[i] + delta); }
        for (int i = 0; i < d.n_x_u; i++) { int k = d.x_u_idx[i]; x_reg[k] += d.x_b_scaling[k] * d.x_b_scaling[k] / (z_bu_inv[i] * s_bu[i] + delta); }

        std::fill(z_reg.begin(), z_reg.end(), 0.0);
        for (int i = 0; i < d.n_h_l; i++) { int k = d.h_l_idx[i]; z_reg[k] += 1.0 / (z_l_inv[k] * s_l[k] + delta); }
        for (int i = 0; i < d.n_h_u; i++) { int k = d.h_u_idx[i]; z_reg[k] += 1.0 / (z_u_inv[k] * s_u[k] + delta); }
        for (int i = 0; i < d.m; i++) { z_reg[i] = 1.0 / z_reg[i]; z_reg_ir[i] = z_reg[i]; }

        double delta_reg = delta;
        if (ir) {
            double max_diag = 0;
            for (int i = 0; i < d.n; i++) max_diag = std::max(max_diag, std::fabs(P_diag[i] + x_reg[i]));
            max_diag = std::max(max_diag, inf_norm(z_reg_ir));
            double reg = st.iterative_refinement_static_regularization_eps + st.iterative_refinement_static_regularization_rel * max_diag;
            delta_reg += reg;
            for (int i = 0; i < d.n; i++) x_reg[i] += reg;
            for (int i = 0; i < d.m; i++) z_reg_ir[i] += reg;
        }
        use_ir = ir;
        return be->factor(delta_reg, x_reg.data(), z_reg_ir.data());
    }

    double norm3(const Vec& a, const Vec& b, const Vec& c) { return std::max(std::max(inf_norm(a), inf_norm(b)), inf_norm(c)); }

    // :507-536 ; err = rhs - K3x3 * lhs
    double refine_error(const Vec& lx, const Vec& ly, const Vec& lz, const Vec& rx, const Vec& ry, const Vec& rz, Vec& ex, Vec& ey, Vec& ez) {
        const int n = (int)lx.size(), p = (int)ly.size(), m = (int)lz.size();
        be->eval_P_x(1.0, lx.data(), ex.data());
        for (int i = 0; i < n; i++) ex[i] += x_reg[i] * lx[i];
        be->eval_A(1.0, 1.0, lx.data(), ly.data(), ey.data(), work_x.data());
        for (int i = 0; i < n; i++) ex[i] += work_x[i];
        for (int i = 0; i < p; i++) ey[i] -= delta * ly[i];
        be->eval_G(1.0, 1.0, lx.data(), lz.data(), ez.data(), work_x.data());
        for (int i = 0; i < n; i++) ex[i] += work_x[i];
        for (int i = 0; i < m; i++) ez[i] -= z_reg[i] * lz[i];
        for (int i = 0; i < n; i++) ex[i] = rx[i] - ex[i];
        for (int i = 0; i < p; i++) ey[i] = ry[i] - ey[i];
        for (int i = 0; i < m; i++) ez[i] = rz[i] - ez[i];
        return norm3(ex, ey, ez);
    }

    static bool all_finite(const Vec& v) { for (double e : v) if (!std::isfinite(e)) return false; return true; }

    // :213-369
    bool solve(const ProblemVectors& d, const Settings& st, const Variables& rhs, Variables& lhs) {
        Vec& lhs_z = work_z;
        std::fill(rhs_z_bar.begin(), rhs_z_bar.end(), 0.0);
        for (int i = 0; i < d.n_h_l; i++) { int k = d.h_l_idx[i];
            rhs_z_bar[k] -= 1.0 / (z_l_inv[k] * s_l[k] + delta) * (rhs.z_l[k] - z_l_inv[k] * rhs.s_l[k]); }
        for (int i = 0; i < d.n_h_u; i++) { int k = d.h_u_idx[i];
            rhs_z_bar[k] += 1.0 / (z_u_inv[k] * s_u[k] + delta) * (rhs.z_u[k] - z_u_inv[k] * rhs.s_u[k]); }
        for (int i = 0; i < d.m; i++) rhs_z_bar[i] *= z_reg[i];

        rhs_x_bar = rhs.x;
        for (int i = 0; i < d.n_x_l; i++) { int k = d.x_l_idx[i];
            rhs_x_bar[k] -= d.x_b_scaling[k] * (rhs.z_bl[i] - z_bl_inv[i] * rhs.s_bl[i]) / (s_bl[i] * z_bl_inv[i] + delta); }
        for (int i = 0; i < d.n_x_u; i++) { int k = d.x_u_idx[i];
            rhs_x_bar[k] += d.x_b_scaling[k] * (rhs.z_bu[i] - z_bu_inv[i] * rhs.s_bu[i]) / (s_bu[i] * z_bu_inv[i] + delta); }

        be->solve(rhs_x_bar.data(), rhs.y.data(), rhs_z_bar.data(), lhs.x.data(), lhs.y.data(), lhs_z.data());
        n_backend_solve++;

        if (use_ir) {  // :256-301
            double rhs_norm = norm3(rhs_x_bar, rhs.y, rhs_z_bar);
            double err = refine_error(lhs.x, lhs.y, lhs_z, rhs_x_bar, rhs.y, rhs_z_bar, err_x, err_y, err_z);
            if (!std::isfinite(err)) return false;
            for (long it = 0; it < st.iterative_refinement_max_iter; it++) {
                if (err <= st.iterative_refinement_eps_abs + st.iterative_refinement_eps_rel * rhs_norm) break;
                double prev = err;
                be->solve(err_x.data(), err_y.data(), err_z.data(), ref_x.data(), ref_y.data(), ref_z.data());
                n_backend_solve++;
                for (size_t i = 0; i < ref_x.size(); i++) ref_x[i] += lhs.x[i];
                for (size_t i = 0; i < ref_y.size(); i++) ref_y[i] += lhs.y[i];
                for (size_t i = 0; i < ref_z.size(); i++) ref_z[i] += lhs_z[i];
                err = refine_error(ref_x, ref_y, ref_z, rhs_x_bar, rhs.y, rhs_z_bar, err_x, err_y, err_z);
                if (!std::isfinite(err)) return false;
                double rate = prev / err;
                if (rate < st.iterative_refinement_min_improvement_rate) {
                    if (rate > 1.0) { std::swap(lhs.x, ref_x); std::swap(lhs.y, ref_y); std::swap(lhs_z, ref_z); }
                    break;
                }
                std::swap(lhs.x, ref_x); std::swap(lhs.y, ref_y); std::swap(lhs_z, ref_z);
            }
        } else {
            if (!all_finite(lhs.x) || !all_finite(lhs.y) || !all_finite(lhs_z)) return false;
        }

        // dual recovery :310-345
        int il = 0, iu = 0;
        for (int i = 0; i < d.m; i++) {
            int kl = il < d.n_h_l ? d.h_l_idx[il] : -1;
            while (kl < i && il < d.n_h_l) { ++il; kl = il < d.n_h_l ? d.h_l_idx[il] : d.m + 1; }
            int ku = iu < d.n_h_u ? d.h_u_idx[iu] : -1;
            while (ku < i && iu < d.n_h_u) { ++iu; ku = iu < d.n_h_u ? d.h_u_idx[iu] : d.m + 1; }
            if (kl == i && ku == i) {
                double rzl = rhs.z_l[i] - z_l_inv[i] * rhs.s_l[i];
                double Wl = 1.0 / (z_l_inv[i] * s_l[i] + delta);
                double rzu = rhs.z_u[i] - z_u_inv[i] * rhs.s_u[i];
                double Wu = 1.0 / (z_u_inv[i] * s_u[i] + delta);
                double rs = Wl * Wu * (rzl + rzu);
                lhs.z_l[i] = -z_reg[i] * (rs + Wl * lhs_z[i]);
                lhs.z_u[i] = -z_reg[i] * (rs - Wu * lhs_z[i]);
                lhs.s_l[i] = z_l_inv[i] * (rhs.s_l[i] - s_l[i] * lhs.z_l[i]);
                lhs.s_u[i] = z_u_inv[i] * (rhs.s_u[i] - s_u[i] * lhs.z_u[i]);
            } else if (kl == i) {
                lhs.z_l[i] = -lhs_z[i]; lhs.z_u[i] = 0;
                lhs.s_l[i] = z_l_inv[i] * (rhs.s_l[i] - s_l[i] * lhs.z_l[i]); lhs.s_u[i] = 0;
            } else if (ku == i) {
                lhs.z_l[i] = 0; lhs.z_u[i] = lhs_z[i];
                lhs.s_l[i] = 0; lhs.s_u[i] = z_u_inv[i] * (rhs.s_u[i] - s_u[i] * lhs.z_u[i]);
            }
            // rows with both sides infinite cannot occur (data.hpp:142-169 rewrites them to [-1,1])
        }
        // box dual recovery :347-366
        for (int i = 0; i < d.n_x_l; i++) { int k = d.x_l_idx[i];
            lhs.z_bl[i] = (-d.x_b_scaling[k] * lhs.x[k] - rhs.z_bl[i] + z_bl_inv[i] * rhs.s_bl[i]) / (s_bl[i] * z_bl_inv[i] + delta); }
        for (int i = 0; i < d.n_x_u; i++) { int k = d.x_u_idx[i];
            lhs.z_bu[i] = (d.x_b_scaling[k] * lhs.x[k] - rhs.z_bu[i] + z_bu_inv[i] * rhs.s_bu[i]) / (s_bu[i] * z_bu_inv[i] + delta); }
        for (int i = 0; i < d.n_x_l; i++) lhs.s_bl[i] = z_bl_inv[i] * (rhs.s_bl[i] - s_bl[i] * lhs.z_bl[i]);
        for (int i = 0; i < d.n_x_u; i++) lhs.s_bu[i] = z_bu_inv[i] * (rhs.s_bu[i] - s_bu[i] * lhs.z_bu[i]);
        return true;
    }

    // :392-425 ; rhs = full 10-block Newton operator * lhs (used by the reference's kkt tests)
    void mul(const ProblemVectors& d, const Variables& lhs, Variables& rhs) {
        const int n = d.n, p = d.p, m = d.m;
        be->eval_P_x(1.0, lhs.x.data(), rhs.x.data());
        for (int i = 0; i < n; i++) rhs.x[i] += rho * lhs.x[i];
        be->eval_A(1.0, 1.0, lhs.x.data(), lhs.y.data(), rhs.y.data(), work_x.data());
        for (int i = 0; i < n; i++) rhs.x[i] += work_x[i];
        for (int i = 0; i < p; i++) rhs.y[i] -= delta * lhs.y[i];
        for (int i = 0; i < m; i++) rhs.s_l[i] = lhs.z_u[i] - lhs.z_l[i];
        be->eval_G(1.0, 1.0, lhs.x.data(), rhs.s_l.data(), rhs.z_u.data(), work_x.data());
        for (int i = 0; i < m; i++) rhs.z_l[i] = -rhs.z_u[i];
        for (int i = 0; i < n; i++) rhs.x[i] += work_x[i];
        for (int i = 0; i < m; i++) {
            rhs.z_l[i] += lhs.s_l[i] - delta * lhs.z_l[i];
            rhs.z_u[i] += lhs.s_u[i] - delta * lhs.z_u[i];
            rhs.s_l[i] = s_l[i] * lhs.z_l[i] + lhs.s_l[i] / z_l_inv[i];
            rhs.s_u[i] = s_u[i] * lhs.z_u[i] + lhs.s_u[i] / z_u_inv[i];
        }
        for (int i = 0; i < d.n_x_l; i++) { int k = d.x_l_idx[i];
            rhs.x[k] -= d.x_b_scaling[k] * lhs.z_bl[i];
            rhs.z_bl[i] = -d.x_b_scaling[k] * lhs.x[k] - delta * lhs.z_bl[i] + lhs.s_bl[i];
            rhs.s_bl[i] = s_bl[i] * lhs.z_bl[i] + lhs.s_bl[i] / z_bl_inv[i]; }
        for (int i = 0; i < d.n_x_u; i++) { int k = d.x_u_idx[i];
            rhs.x[k] += d.x_b_scaling[k] * lhs.z_bu[i];
            rhs.z_bu[i] = d.x_b_scaling[k] * lhs.x[k] - delta * lhs.z_bu[i] + lhs.s_bu[i];
            rhs.s_bu[i] = s_bu[i] * lhs.z_bu[i] + lhs.s_bu[i] / z_bu_inv[i]; }
    }
};

// solver.hpp:34-1260 (SolverBase), matrix-type agnostic.
struct IPSolver {
    Settings st;
    Info info;
    Variables res;                 // m_result (iterate) -- named `it` below
    ProblemVectors d;
    std::unique_ptr<QPMatrices> M;
    Ruiz pre;
    KKTSystem kkt;
    bool first_run = true, setup_done = false, ir_on = false;
    Variables it;                  // current iterate (m_result)
    Variables rnr;                 // res_nr (BasicVariables part used)
    Variables r;                   // res
    Variables step;
    Variables prox;                // prox_vars
    std::vector<double> trace;     // per-iteration (rho, delta, mu, primal_step, dual_step, primal_res, dual_res)

    // solver.hpp:151-216 after the matrices were stored by the caller
    bool finish_setup(const double* c, const double* b, const double* h_l, const double* h_u, const double* x_l, const double* x_u) {
        for (int i = 0; i < d.n; i++) d.c[i] = c[i];
        if (b) for (int i = 0; i < d.p; i++) d.b[i] = b[i];
        d.set_h_l(h_l); d.set_h_u(h_u);
        disable_inf_constraints();
        d.set_x_l(x_l); d.set_x_u(x_u);
        init_workspace();
        pre.init(d);
        pre.scale_data(d, *M, false, st.preconditioner_scale_cost, st.preconditioner_iter);
        if (!kkt.init(d, *M, st)) { setup_done = false; return false; }
        first_run = true; setup_done = true;
        return true;
    }

    // dense/data.hpp:142-169
    void disable_inf_constraints() {
        bool any = false;
        for (int i = 0; i < d.m; i++) {
            if (d.h_l[i] <= -kInf && d.h_u[i] >= kInf) { M->set_G_row_zero(i); d.h_l[i] = -1; d.h_u[i] = 1; any = true; }
        }
        if (any) { Vec hl = d.h_l, hu = d.h_u; d.set_h_l(hl.data()); d.set_h_u(hu.data()); }
    }

    void init_workspace() {  // solver.hpp:361-377
        it.resize(d.n, d.p, d.m);
        info = Info();
        info.rho = st.rho_init; info.delta = st.delta_init;
        rnr.resize(d.n, d.p, d.m); r.resize(d.n, d.p, d.m); step.resize(d.n, d.p, d.m); prox.resize(d.n, d.p, d.m);
    }

    // solver.hpp:218-308.  The caller overwrites the (unscaled) matrices between begin_update and end_update.
    void begin_update() { pre.unscale_data(d, *M); }
    void end_update(int options, const double* c, const double* b, const double* h_l, const double* h_u, const double* x_l, const double* x_u) {
        if (c) for (int i = 0; i < d.n; i++) d.c[i] = c[i];
        if (b) for (int i = 0; i < d.p; i++) d.b[i] = b[i];
        if (h_l) d.set_h_l(h_l);
        if (h_u) d.set_h_u(h_u);
        if (h_l || h_u) disable_inf_constraints();
        if (x_l) d.set_x_l(x_l);
        if (x_u) d.set_x_u(x_u);
        bool reuse = st.preconditioner_reuse_on_update;
        if (options == UPDATE_NONE) reuse = true;
        pre.scale_data(d, *M, reuse, st.preconditioner_scale_cost, st.preconditioner_iter);
        kkt.update_data(*M, options);
    }

    double calc_mu() const {  // :884-891
        double s = dot(it.s_l.data(), it.z_l.data(), d.m) + dot(it.s_u.data(), it.z_u.data(), d.m)
                 + dot(it.s_bl.data(), it.z_bl.data(), d.n_x_l) + dot(it.s_bu.data(), it.z_bu.data(), d.n_x_u);
        return s / double(d.n_h_l + d.n_h_u + d.n_x_l + d.n_x_u);
    }

    void calc_step(double& as, double& az) const {  // :893-958
        as = 1; az = 1;
        for (int i = 0; i < d.m; i++) {
            if (step.s_l[i] < 0) as = std::min(as, -it.s_l[i] / step.s_l[i]);
            if (step.s_u[i] < 0) as = std::min(as, -it.s_u[i] / step.s_u[i]);
            if (step.z_l[i] < 0) az = std::min(az, -it.z_l[i] / step.z_l[i]);
            if (step.z_u[i] < 0) az = std::min(az, -it.z_u[i] / step.z_u[i]);
        }
        for (int i = 0; i < d.n_x_l; i++) {
            if (step.s_bl[i] < 0) as = std::min(as, -it.s_bl[i] / step.s_bl[i]);
            if (step.z_bl[i] < 0) az = std::min(az, -it.z_bl[i] / step.z_bl[i]);
        }
        for (int i = 0; i < d.n_x_u; i++) {
            if (step.s_bu[i] < 0) as = std::min(as, -it.s_bu[i] / step.s_bu[i]);
            if (step.z_bu[i] < 0) az = std::min(az, -it.z_bu[i] / step.z_bu[i]);
        }
    }

    bool factor_with_retry(bool in_loop, bool& reg_changed) {  // :446-465 / :688-708
        while (!kkt.update_scalings_and_factor(d, st, ir_on, info.rho, info.delta, it)) {
            info.n_factor++;
            if (!ir_on) { ir_on = true; continue; }
            if (info.factor_retires < st.max_factor_retires) {
                info.delta *= 100; info.rho *= 100; info.factor_retires++;
                info.reg_limit = std::min(10 * info.reg_limit, st.eps_abs);
                if (in_loop) reg_changed = true;
                continue;
            }
            info.status = NUMERICS;
            return false;
        }
        info.n_factor++;
        info.factor_retires = 0;
        return true;
    }

    int solve() {  // :69-148
        int s = solve_impl();
        unscale_results();
        restore_dual();
        info.n_backend_solve = kkt.n_backend_solve;
        first_run = false;
        return s;
    }

    int solve_impl() {  // :379-882
        const int n = d.n, p = d.p, m = d.m;
        if (!setup_done) { info.status = UNSOLVED; return info.status; }
        if (!st.verify()) { info.status = INVALID_SETTINGS; return info.status; }
        trace.clear();
        kkt.n_backend_solve = 0;
        info.n_factor = info.n_solve = 0;
        info.status = UNSOLVED; info.iter = 0; info.reg_limit = st.reg_lower_limit;
        info.factor_retires = 0; info.no_primal_update = 0; info.no_dual_update = 0;
        info.mu = 0; info.primal_step = 0; info.dual_step = 0;
        info.rho = st.rho_init; info.delta = st.delta_init;

        std::fill(it.s_l.begin(), it.s_l.end(), 0.0); std::fill(it.s_u.begin(), it.s_u.end(), 0.0);
        std::fill(it.z_l.begin(), it.z_l.end(), 0.0); std::fill(it.z_u.begin(), it.z_u.end(), 0.0);
        for (int i = 0; i < d.n_h_l; i++) { it.s_l[d.h_l_idx[i]] = 1; it.z_l[d.h_l_idx[i]] = 1; }
        for (int i = 0; i < d.n_h_u; i++) { it.s_u[d.h_u_idx[i]] = 1; it.z_u[d.h_u_idx[i]] = 1; }
        for (int i = 0; i < d.n_x_l; i++) { it.s_bl[i] = 1; it.z_bl[i] = 1; }
        for (int i = 0; i < d.n_x_u; i++) { it.s_bu[i] = 1; it.z_bu[i] = 1; }

        ir_on = st.iterative_refinement_always_enabled;
        bool dummy = false;
        if (!factor_with_retry(false, dummy)) return info.status;

        for (int i = 0; i < n; i++) r.x[i] = -d.c[i];
        for (int i = 0; i < p; i++) r.y[i] = d.b[i];
        for (int i = 0; i < m; i++) { r.z_l[i] = -d.h_l[i]; r.z_u[i] = d.h_u[i]; r.s_l[i] = 0; r.s_u[i] = 0; }
        for (int i = 0; i < n; i++) { r.z_bl[i] = -d.x_l[i]; r.z_bu[i] = d.x_u[i]; r.s_bl[i] = 0; r.s_bu[i] = 0; }
        kkt.solve(d, st, r, it); info.n_solve++;

        const int nl = d.n_x_l, nu = d.n_x_u;
        if (m + nl + nu > 0) {  // :504-570
            double ds = 0, dz = 0;
            auto minc = [](const Vec& v, int k) { double r_ = v[0]; for (int i = 1; i < k; i++) r_ = std::min(r_, v[i]); return r_; };
            if (m > 0) { ds = std::max(ds, -minc(it.s_l, m)); ds = std::max(ds, -minc(it.s_u, m)); }
            if (nl > 0) ds = std::max(ds, -minc(it.s_bl, nl));
            if (nu > 0) ds = std::max(ds, -minc(it.s_bu, nu));
            if (m > 0) { dz = std::max(dz, -minc(it.z_l, m)); dz = std::max(dz, -minc(it.z_u, m)); }
            if (nl > 0) dz = std::max(dz, -minc(it.z_bl, nl));
            if (nu > 0) dz = std::max(dz, -minc(it.z_bu, nu));
            for (int i = 0; i < d.n_h_l; i++) { int k = d.h_l_idx[i]; it.s_l[k] += ds; it.z_l[k] += dz; }
            for (int i = 0; i < d.n_h_u; i++) { int k = d.h_u_idx[i]; it.s_u[k] += ds; it.z_u[k] += dz; }
            for (int i = 0; i < nl; i++) { it.s_bl[i] += ds; it.z_bl[i] += dz; }
            for (int i = 0; i < nu; i++) { it.s_bu[i] += ds; it.z_bu[i] += dz; }
            info.mu = std::max(calc_mu(), 1e-10);
            auto fix = [&](double& z, double& s) { double c_ = z - dz; z = (c_ + std::sqrt(c_ * c_ + 4 * info.mu)) / 2; s = z - c_; };
            for (int i = 0; i < d.n_h_l; i++) { int k = d.h_l_idx[i]; fix(it.z_l[k], it.s_l[k]); }
            for (int i = 0; i < d.n_h_u; i++) { int k = d.h_u_idx[i]; fix(it.z_u[k], it.s_u[k]); }
            for (int i = 0; i < nl; i++) fix(it.z_bl[i], it.s_bl[i]);
            for (int i = 0; i < nu; i++) fix(it.z_bu[i], it.s_bu[i]);
            info.mu = calc_mu();
        }
        prox.x = it.x; prox.y = it.y; prox.z_l = it.z_l; prox.z_u = it.z_u;
        for (int i = 0; i < nl; i++) prox.z_bl[i] = it.z_bl[i];
        for (int i = 0; i < nu; i++) prox.z_bu[i] = it.z_bu[i];

        const double eps = std::numeric_limits<double>::epsilon();
        while (info.iter < st.max_iter) {
            if (info.iter == 0) { update_residuals_nr(); info.prev_primal_res = info.primal_res; info.prev_dual_res = info.dual_res; }
            trace.insert(trace.end(), {info.rho, info.delta, info.mu, info.primal_step, info.dual_step, info.primal_res, info.dual_res,
                                       info.primal_obj, info.dual_obj, info.duality_gap});
            if (st.verbose)
                printf("%3ld   % .5e   % .5e   %.5e   %.5e   %.5e   %.3e   %.3e   %.3e   %.4f   %.4f\n", info.iter, info.primal_obj,
                       info.dual_obj, info.duality_gap, info.primal_res, info.dual_res, info.rho, info.delta, info.mu, info.primal_step, info.dual_step);

            if ((info.primal_res < st.eps_abs || info.primal_res_rel < st.eps_rel) &&
                (info.dual_res < st.eps_abs || info.dual_res_rel < st.eps_rel) &&
                (!st.check_duality_gap || info.duality_gap < st.eps_duality_gap_abs || info.duality_gap_rel < st.eps_duality_gap_rel)) {
                info.status = SOLVED; return info.status;
            }
            update_residuals_r();
            if (info.no_dual_update > std::min(5L, st.reg_finetune_dual_update_threshold) &&
                info.primal_prox_inf > st.infeasibility_threshold &&
                (info.primal_res_reg < st.eps_abs || info.primal_res_reg_rel < st.eps_rel)) {
                info.status = PRIMAL_INFEASIBLE; return info.status;
            }
            if (info.no_primal_update > std::min(5L, st.reg_finetune_primal_update_threshold) &&
                info.dual_prox_inf > st.infeasibility_threshold &&
                (info.dual_res_reg < st.eps_abs || info.dual_res_reg_rel < st.eps_rel)) {
                info.status = DUAL_INFEASIBLE; return info.status;
            }
            info.iter++;

            bool shifted = false;  // :634-666
            for (int i = 0; i < d.n_h_l; i++) { int k = d.h_l_idx[i]; if (it.z_l[k] < eps) { it.z_l[k] += eps; shifted = true; } }
            for (int i = 0; i < d.n_h_u; i++) { int k = d.h_u_idx[i]; if (it.z_u[k] < eps) { it.z_u[k] += eps; shifted = true; } }
            if (nl > 0) { double mn = it.z_bl[0]; for (int i = 1; i < nl; i++) mn = std::min(mn, it.z_bl[i]);
                if (mn < eps) { for (int i = 0; i < nl; i++) it.z_bl[i] += eps; shifted = true; } }
            if (nu > 0) { double mn = it.z_bu[0]; for (int i = 1; i < nu; i++) mn = std::min(mn, it.z_bu[i]);
                if (mn < eps) { for (int i = 0; i < nu; i++) it.z_bu[i] += eps; shifted = true; } }
            if (shifted) info.mu = calc_mu();

            if ((info.no_primal_update > st.reg_finetune_primal_update_threshold && info.rho == info.reg_limit && info.reg_limit != st.reg_finetune_lower_limit) ||
                (info.no_dual_update > st.reg_finetune_dual_update_threshold && info.delta == info.reg_limit && info.reg_limit != st.reg_finetune_lower_limit)) {
                if (info.dual_prox_inf < st.infeasibility_threshold && info.primal_prox_inf < st.infeasibility_threshold) {
                    info.reg_limit = st.reg_finetune_lower_limit; info.no_primal_update = 0; info.no_dual_update = 0;
                }
            }

            bool reg_changed = false;
            if (!factor_with_retry(true, reg_changed)) return info.status;
            if (reg_changed) update_residuals_r();

            if (m + nl + nu > 0) {
                for (int i = 0; i < m; i++) { r.s_l[i] = -it.s_l[i] * it.z_l[i]; r.s_u[i] = -it.s_u[i] * it.z_u[i]; }
                for (int i = 0; i < nl; i++) r.s_bl[i] = -it.s_bl[i] * it.z_bl[i];
                for (int i = 0; i < nu; i++) r.s_bu[i] = -it.s_bu[i] * it.z_bu[i];
                kkt.solve(d, st, r, step); info.n_solve++;
                double as, az; calc_step(as, az);
                as *= st.tau; az *= st.tau;
                double sg = 0, acc;
                acc = 0; for (int i = 0; i < m; i++) acc += (it.s_l[i] + as * step.s_l[i]) * (it.z_l[i] + az * step.z_l[i]); sg = acc;
                acc = 0; for (int i = 0; i < m; i++) acc += (it.s_u[i] + as * step.s_u[i]) * (it.z_u[i] + az * step.z_u[i]); sg += acc;
                acc = 0; for (int i = 0; i < nl; i++) acc += (it.s_bl[i] + as * step.s_bl[i]) * (it.z_bl[i] + az * step.z_bl[i]); sg += acc;
                acc = 0; for (int i = 0; i < nu; i++) acc += (it.s_bu[i] + as * step.s_bu[i]) * (it.z_bu[i] + az * step.z_bu[i]); sg += acc;
                sg /= (info.mu * double(d.n_h_l + d.n_h_u + nl + nu));
                sg = std::max(0.0, std::min(1.0, sg));
                info.sigma = sg * sg * sg;
                const double sm = info.sigma * info.mu;
                for (int i = 0; i < m; i++) { r.s_l[i] += -step.s_l[i] * step.z_l[i] + sm; r.s_u[i] += -step.s_u[i] * step.z_u[i] + sm; }
                for (int i = 0; i < nl; i++) r.s_bl[i] += -step.s_bl[i] * step.z_bl[i] + sm;
                for (int i = 0; i < nu; i++) r.s_bu[i] += -step.s_bu[i] * step.z_bu[i] + sm;
                kkt.solve(d, st, r, step); info.n_solve++;
                calc_step(as, az);
                info.primal_step = as * st.tau; info.dual_step = az * st.tau;
                for (int i = 0; i < n; i++) it.x[i] += info.primal_step * step.x[i];
                for (int i = 0; i < p; i++) it.y[i] += info.dual_step * step.y[i];
                for (int i = 0; i < m; i++) { it.z_l[i] += info.dual_step * step.z_l[i]; it.z_u[i] += info.dual_step * step.z_u[i]; }
                for (int i = 0; i < nl; i++) it.z_bl[i] += info.dual_step * step.z_bl[i];
                for (int i = 0; i < nu; i++) it.z_bu[i] += info.dual_step * step.z_bu[i];
                for (int i = 0; i < m; i++) { it.s_l[i] += info.primal_step * step.s_l[i]; it.s_u[i] += info.primal_step * step.s_u[i]; }
                for (int i = 0; i < nl; i++) it.s_bl[i] += info.primal_step * step.s_bl[i];
                for (int i = 0; i < nu; i++) it.s_bu[i] += info.primal_step * step.s_bu[i];
                double mu_prev = info.mu;
                info.mu = calc_mu();
                double mu_rate = std::max(0.0, (mu_prev - info.mu) / mu_prev);
                update_residuals_nr();
                if (info.dual_res < 0.95 * info.prev_dual_res || (info.dual_res < st.eps_abs || info.dual_res_rel < st.eps_rel) ||
                    (info.rho == st.reg_finetune_lower_limit && info.dual_prox_inf < st.infeasibility_threshold)) {
                    prox.x = it.x;
                    info.rho = std::max(info.reg_limit, (1.0 - mu_rate) * info.rho);
                } else {
                    info.no_primal_update++;
                    if (info.iter < 5 || info.dual_prox_inf < st.infeasibility_threshold)
                        info.rho = std::max(info.reg_limit, (1.0 - 0.666 * mu_rate) * info.rho);
                }
                if (info.primal_res < 0.95 * info.prev_primal_res || (info.primal_res < st.eps_abs || info.primal_res_rel < st.eps_rel) ||
                    (info.delta == st.reg_finetune_lower_limit && info.primal_prox_inf < st.infeasibility_threshold)) {
                    prox.y = it.y; prox.z_l = it.z_l; prox.z_u = it.z_u;
                    for (int i = 0; i < nl; i++) prox.z_bl[i] = it.z_bl[i];
                    for (int i = 0; i < nu; i++) prox.z_bu[i] = it.z_bu[i];
                    info.delta = std::max(info.reg_limit, (1.0 - mu_rate) * info.delta);
                } else {
                    info.no_dual_update++;
                    if (info.iter < 5 || info.primal_prox_inf < st.infeasibility_threshold)
                        info.delta = std::max(info.reg_limit, (1.0 - 0.666 * mu_rate) * info.delta);
                }
            } else {  // :831-877
                kkt.solve(d, st, r, step); info.n_solve++;
                info.primal_step = 1; info.dual_step = 1;
                for (int i = 0; i < n; i++) it.x[i] += info.primal_step * step.x[i];
                for (int i = 0; i < p; i++) it.y[i] += info.dual_step * step.y[i];
                update_residuals_nr();
                if (info.dual_res < 0.95 * info.prev_dual_res || (info.dual_res < st.eps_abs || info.dual_res_rel < st.eps_rel)) {
                    prox.x = it.x; info.rho = std::max(info.reg_limit, 0.1 * info.rho);
                } else {
                    info.no_primal_update++;
                    if (info.iter < 5 || info.dual_prox_inf < st.infeasibility_threshold) info.rho = std::max(info.reg_limit, 0.5 * info.rho);
                }
                if (info.primal_res < 0.95 * info.prev_primal_res || (info.primal_res < st.eps_abs || info.primal_res_rel < st.eps_rel)) {
                    prox.y = it.y; info.delta = std::max(info.reg_limit, 0.1 * info.delta);
                } else {
                    info.no_dual_update++;
                    if (info.iter < 5 || info.primal_prox_inf < st.infeasibility_threshold) info.delta = std::max(info.reg_limit, 0.5 * info.delta);
                }
            }
        }
        info.status = MAX_ITER_REACHED;
        return info.status;
    }

    void update_residuals_nr() {  // :960-1105
        const int n = d.n, p = d.p, m = d.m;
        Vec& wx = step.x; Vec& wz = step.z_l;
        kkt.be->eval_A(-1.0, 1.0, it.x.data(), it.y.data(), rnr.y.data(), wx.data());
        for (int i = 0; i < m; i++) wz[i] = it.z_u[i] - it.z_l[i];
        Vec& wx2 = rnr.x;
        kkt.be->eval_G(1.0, 1.0, it.x.data(), wz.data(), rnr.z_l.data(), wx2.data());
        for (int i = 0; i < m; i++) rnr.z_u[i] = -rnr.z_l[i];
        for (int i = 0; i < n; i++) wx[i] += wx2[i];

        kkt.be->eval_P_x(-1.0, it.x.data(), rnr.x.data());
        double dual_rel = 0; for (int i = 0; i < n; i++) dual_rel = std::max(dual_rel, std::fabs(pre.us_dres(rnr.x[i], i)));

        double tmp = -dot(it.x.data(), rnr.x.data(), n);
        info.primal_obj = 0.5 * tmp; info.dual_obj = -0.5 * tmp;
        double gap_rel = pre.unscale_cost(std::fabs(tmp));
        tmp = dot(d.c.data(), it.x.data(), n); info.primal_obj += tmp; gap_rel = std::max(gap_rel, pre.unscale_cost(std::fabs(tmp)));
        tmp = dot(d.b.data(), it.y.data(), p); info.dual_obj -= tmp; gap_rel = std::max(gap_rel, pre.unscale_cost(std::fabs(tmp)));
        tmp = -dot(d.h_l.data(), it.z_l.data(), m); info.dual_obj -= tmp; gap_rel = std::max(gap_rel, pre.unscale_cost(std::fabs(tmp)));
        tmp = dot(d.h_u.data(), it.z_u.data(), m); info.dual_obj -= tmp; gap_rel = std::max(gap_rel, pre.unscale_cost(std::fabs(tmp)));
        tmp = -dot(d.x_l.data(), it.z_bl.data(), d.n_x_l); info.dual_obj -= tmp; gap_rel = std::max(gap_rel, pre.unscale_cost(std::fabs(tmp)));
        tmp = dot(d.x_u.data(), it.z_bu.data(), d.n_x_u); info.dual_obj -= tmp; gap_rel = std::max(gap_rel, pre.unscale_cost(std::fabs(tmp)));
        info.duality_gap = std::fabs(info.primal_obj - info.dual_obj);
        info.primal_obj = pre.unscale_cost(info.primal_obj);
        info.dual_obj = pre.unscale_cost(info.dual_obj);
        info.duality_gap = pre.unscale_cost(info.duality_gap);
        info.duality_gap_rel = info.duality_gap / std::max(1.0, gap_rel);

        for (int i = 0; i < n; i++) rnr.x[i] -= d.c[i];
        for (int i = 0; i < n; i++) dual_rel = std::max(dual_rel, std::fabs(pre.us_dres(d.c[i], i)));
        for (int i = 0; i < d.n_x_l; i++) { int k = d.x_l_idx[i]; wx[k] -= d.x_b_scaling[k] * it.z_bl[i]; }
        for (int i = 0; i < d.n_x_u; i++) { int k = d.x_u_idx[i]; wx[k] += d.x_b_scaling[k] * it.z_bu[i]; }
        for (int i = 0; i < n; i++) dual_rel = std::max(dual_rel, std::fabs(pre.us_dres(wx[i], i)));
        for (int i = 0; i < n; i++) rnr.x[i] -= wx[i];

        double prim_rel = 0; for (int i = 0; i < p; i++) prim_rel = std::max(prim_rel, std::fabs(pre.us_pres_eq(rnr.y[i], i)));
        for (int i = 0; i < p; i++) rnr.y[i] += d.b[i];
        for (int i = 0; i < p; i++) prim_rel = std::max(prim_rel, std::fabs(pre.us_pres_eq(d.b[i], i)));

        // NOTE: signed values (no abs) enter the max for the inequality/box terms, as in the reference :1047-1093
        int i = 0;
        for (int ii = 0; ii < d.n_h_l; ii++) {
            int k = d.h_l_idx[ii];
            while (i < k) rnr.z_l[i++] = 0;
            prim_rel = std::max(prim_rel, pre.us_pres_ineq(rnr.z_l[i], i));
            rnr.z_l[i] += -d.h_l[i] - it.s_l[i];
            prim_rel = std::max(prim_rel, pre.us_pres_ineq(d.h_l[i], i));
            prim_rel = std::max(prim_rel, pre.us_pres_ineq(it.s_l[i], i));
            i++;
        }
        while (i < m) rnr.z_l[i++] = 0;
        i = 0;
        for (int ii = 0; ii < d.n_h_u; ii++) {
            int k = d.h_u_idx[ii];
            while (i < k) rnr.z_u[i++] = 0;
            prim_rel = std::max(prim_rel, pre.us_pres_ineq(rnr.z_u[i], i));
            rnr.z_u[i] += d.h_u[i] - it.s_u[i];
            prim_rel = std::max(prim_rel, pre.us_pres_ineq(d.h_u[i], i));
            prim_rel = std::max(prim_rel, pre.us_pres_ineq(it.s_u[i], i));
            i++;
        }
        while (i < m) rnr.z_u[i++] = 0;
        for (i = 0; i < d.n_x_l; i++) { int k = d.x_l_idx[i];
            rnr.z_bl[i] = d.x_b_scaling[k] * it.x[k];
            prim_rel = std::max(prim_rel, pre.us_pres_b(rnr.z_bl[i], k));
            prim_rel = std::max(prim_rel, pre.us_pres_b(d.x_l[i], k));
            prim_rel = std::max(prim_rel, pre.us_pres_b(it.s_bl[i], k)); }
        for (i = 0; i < d.n_x_l; i++) rnr.z_bl[i] += -d.x_l[i] - it.s_bl[i];
        for (i = 0; i < d.n_x_u; i++) { int k = d.x_u_idx[i];
            rnr.z_bu[i] = -d.x_b_scaling[k] * it.x[k];
            prim_rel = std::max(prim_rel, pre.us_pres_b(rnr.z_bu[i], k));
            prim_rel = std::max(prim_rel, pre.us_pres_b(d.x_u[i], k));
            prim_rel = std::max(prim_rel, pre.us_pres_b(it.s_bu[i], k)); }
        for (i = 0; i < d.n_x_u; i++) rnr.z_bu[i] += d.x_u[i] - it.s_bu[i];

        info.prev_primal_res = info.primal_res; info.prev_dual_res = info.dual_res;
        info.primal_res = primal_res(rnr);
        info.primal_res_rel = info.primal_res / std::max(1.0, prim_rel);
        info.dual_res = dual_res(rnr);
        info.dual_res_rel = info.dual_res / std::max(1.0, dual_rel);
    }

    double primal_res(const Variables& v) const {  // :1130-1164
        double r_ = 0;
        for (int i = 0; i < d.p; i++) r_ = std::max(r_, std::fabs(pre.us_pres_eq(v.y[i], i)));
        for (int i = 0; i < d.m; i++) r_ = std::max(r_, std::fabs(pre.us_pres_ineq(v.z_l[i], i)));
        for (int i = 0; i < d.m; i++) r_ = std::max(r_, std::fabs(pre.us_pres_ineq(v.z_u[i], i)));
        for (int i = 0; i < d.n_x_l; i++) r_ = std::max(r_, pre.us_pres_b(v.z_bl[i], d.x_l_idx[i]));
        for (int i = 0; i < d.n_x_u; i++) r_ = std::max(r_, pre.us_pres_b(v.z_bu[i], d.x_u_idx[i]));
        return r_;
    }
    double dual_res(const Variables& v) const {  // :1184-1196
        double r_ = 0; for (int i = 0; i < d.n; i++) r_ = std::max(r_, std::fabs(pre.us_dres(v.x[i], i))); return r_;
    }

    void update_residuals_r() {  // :1107-1128
        const int n = d.n, p = d.p, m = d.m;
        for (int i = 0; i < n; i++) r.x[i] = rnr.x[i] - info.rho * (it.x[i] - prox.x[i]);
        for (int i = 0; i < p; i++) r.y[i] = rnr.y[i] - info.delta * (prox.y[i] - it.y[i]);
        for (int i = 0; i < m; i++) { r.z_l[i] = rnr.z_l[i] - info.delta * (prox.z_l[i] - it.z_l[i]); r.z_u[i] = rnr.z_u[i] - info.delta * (prox.z_u[i] - it.z_u[i]); }
        for (int i = 0; i < d.n_x_l; i++) r.z_bl[i] = rnr.z_bl[i] - info.delta * (prox.z_bl[i] - it.z_bl[i]);
        for (int i = 0; i < d.n_x_u; i++) r.z_bu[i] = rnr.z_bu[i] - info.delta * (prox.z_bu[i] - it.z_bu[i]);
        double ps = info.primal_res_rel > 0 ? info.primal_res / info.primal_res_rel : 1.0;
        double dsn = info.dual_res_rel > 0 ? info.dual_res / info.dual_res_rel : 1.0;
        info.primal_res_reg = primal_res(r); info.primal_res_reg_rel = info.primal_res_reg / ps;
        info.dual_res_reg = dual_res(r); info.dual_res_reg_rel = info.dual_res_reg / dsn;
        // primal_prox_inf :1166-1182, dual_prox_inf :1198-1203
        double pi = 0;
        for (int i = 0; i < p; i++) pi = std::max(pi, std::fabs(pre.us_dual_eq(prox.y[i] - it.y[i], i)));
        for (int i = 0; i < m; i++) pi = std::max(pi, std::fabs(pre.us_dual_ineq(prox.z_l[i] - it.z_l[i], i)));
        for (int i = 0; i < m; i++) pi = std::max(pi, std::fabs(pre.us_dual_ineq(prox.z_u[i] - it.z_u[i], i)));
        for (int i = 0; i < d.n_x_l; i++) pi = std::max(pi, pre.us_dual_b(prox.z_bl[i] - it.z_bl[i], d.x_l_idx[i]));
        for (int i = 0; i < d.n_x_u; i++) pi = std::max(pi, pre.us_dual_b(prox.z_bu[i] - it.z_bu[i], d.x_u_idx[i]));
        info.primal_prox_inf = pi * info.delta;
        double di = 0; for (int i = 0; i < n; i++) di = std::max(di, std::fabs(pre.us_primal(it.x[i] - prox.x[i], i)));
        info.dual_prox_inf = di * info.rho;
    }

    void unscale_results() {  // :1205-1227
        if (pre.identity) return;
        for (int i = 0; i < d.n; i++) it.x[i] = pre.us_primal(it.x[i], i);
        for (int i = 0; i < d.p; i++) it.y[i] = pre.us_dual_eq(it.y[i], i);
        for (int i = 0; i < d.m; i++) { it.z_l[i] = pre.us_dual_ineq(it.z_l[i], i); it.z_u[i] = pre.us_dual_ineq(it.z_u[i], i);
            it.s_l[i] = pre.us_slack_ineq(it.s_l[i], i); it.s_u[i] = pre.us_slack_ineq(it.s_u[i], i); }
        for (int i = 0; i < d.n_x_l; i++) { int k = d.x_l_idx[i]; it.z_bl[i] = pre.us_dual_b(it.z_bl[i], k); it.s_bl[i] = pre.us_slack_b(it.s_bl[i], k); }
        for (int i = 0; i < d.n_x_u; i++) { int k = d.x_u_idx[i]; it.z_bu[i] = pre.us_dual_b(it.z_bu[i], k); it.s_bu[i] = pre.us_slack_b(it.s_bu[i], k); }
    }
    void restore_dual() {  // :1229-1259
        for (int i = 0; i < d.m; i++) { if (it.z_l[i] == 0) it.s_l[i] = kInf; if (it.z_u[i] == 0) it.s_u[i] = kInf; }
        for (int i = d.n_x_l; i < d.n; i++) { it.z_bl[i] = 0; it.s_bl[i] = kInf; }
        for (int i = d.n_x_u; i < d.n; i++) { it.z_bu[i] = 0; it.s_bu[i] = kInf; }
        for (int i = d.n_x_l - 1; i >= 0; i--) { int k = d.x_l_idx[i]; std::swap(it.z_bl[i], it.z_bl[k]); std::swap(it.s_bl[i], it.s_bl[k]); }
        for (int i = d.n_x_u - 1; i >= 0; i--) { int k = d.x_u_idx[i]; std::swap(it.z_bu[i], it.z_bu[k]); std::swap(it.s_bu[i], it.s_bu[k]); }
    }
};

}  // namespace oracle
