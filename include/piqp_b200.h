/*
 * piqp_b200.h -- C-ABI of libpiqp_b200.so: a B200 (sm_100a) KKT factorise-and-solve backend for PIQP.
 *
 * Two layers, both plain C (no C++/torch types cross this boundary, nothing throws across it):
 *
 *  (1) b200kkt_*  -- ONE KKT backend instance.  This is the drop-in for the reference's plugin
 *      interface `piqp::KKTSolverBase<T,I,MatrixType>` (include/piqp/kkt_solver_base.hpp:21-44):
 *      every entry point below names the virtual it replaces.  Host pointers in, host pointers out;
 *      the H2D/D2H copies of the O(n+p+m) vectors happen inside the call.  A maintainer binds it with
 *      the adapter class shown in INTEGRATION.md (`b200::DenseKKT : KKTSolverBase<double,int,PIQP_DENSE>`
 *      etc.) and three new `KKTSolver` enum values in `KKTSystem::init_kkt_solver`
 *      (include/piqp/kkt_system.hpp:455-497).
 *
 *  (2) b200qp_*   -- a BATCH of independent QPs solved with a device-resident interior-point loop
 *      (the reference's `SolverBase::solve_impl`, include/piqp/solver.hpp:379-882, runs per instance on
 *      the GPU; one tiny flag read-back per iteration).  Mirrors the reference's public C API
 *      (interfaces/c/include/piqp.h:21-43: piqp_setup_dense / piqp_update_dense / piqp_solve) with a
 *      leading `batch` dimension; settings and info structs are layout-identical to
 *      `piqp_settings` / `piqp_info` (interfaces/c/include/piqp_typedef.h:75-159).
 *
 * Conventions: fp64 values, int32 indices (include/piqp/common.hpp:38-39).  Dense matrices are
 * row-major like the reference C API (piqp_typedef.h:41-54), i.e. A (p x n, row-major) is exactly
 * AT (n x p, column-major) = the layout `dense::Data::AT` has (include/piqp/dense/data.hpp:30-32).
 * Infinite bounds: |v| >= 1e30 (PIQP_INF, include/piqp/fwd.hpp:54) or IEEE inf.
 * Return codes: 0 = ok (or 1/0 for b200kkt_factor), negative = error (see B200_E_*).
 */
#ifndef PIQP_B200_H
#define PIQP_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define B200_OK 0
#define B200_E_INVALID (-1)   /* bad argument / dimension mismatch              */
#define B200_E_CUDA (-2)      /* CUDA runtime error (see b200_last_error())      */
#define B200_E_NOT_SETUP (-3) /* call order violated                             */
#define B200_E_UNSUPPORTED (-4)

#define B200_INF 1e30
#define B200_MAX_BATCH 65535 /* instances per b200qp handle (larger batches: several handles) */

/* kkt_fwd.hpp:23-29 (KKTUpdateOptions) */
#define B200_KKT_UPDATE_NONE 0
#define B200_KKT_UPDATE_P 1
#define B200_KKT_UPDATE_A 2
#define B200_KKT_UPDATE_G 4

const char* b200_last_error(void);
/* number of kernels launched by this library in this process so far (bench.py reports the delta) */
unsigned long long b200_kernel_launch_count(void);
/* debugging aid: with B200_TIMELINE=1 in the environment every kernel launch is followed by a one-thread kernel that stamps
 * %globaltimer; this writes "index <TAB> kernel <TAB> ns" per stamp in execution order (graph replays included), starts a new
 * timeline and returns the number of stamps written, 0 when the timeline is off, -1 on error                                  */
int b200_timeline_dump(const char* path);
int b200_device_count(void);

/* ------------------------------------------------------------------------------------------------
 * (1) single-instance KKT backend  ==  piqp::KKTSolverBase
 * ------------------------------------------------------------------------------------------------ */
typedef struct b200kkt_handle b200kkt_handle;

/* replaces dense::KKT<T>::KKT(const Data&)            (include/piqp/dense/kkt.hpp:39-55)
 * P_utri: n x n column-major, upper triangle used; AT: n x p column-major; GT: n x m column-major
 * (the already Ruiz-scaled members of dense::Data).                                              */
int b200kkt_dense_create(b200kkt_handle** out, int n, int p, int m,
                         const double* P_utri, const double* AT, const double* GT, int device);

/* replaces sparse::KKT<T,I,Mode>::KKT(const Data&) (include/piqp/sparse/kkt.hpp:51-70)
 * CSC of P_utri (upper), AT (n x p), GT (n x m).  mode = KKTMode (kkt_fwd.hpp:15-21): 0 FULL (sparse_ldlt, kkt_full.hpp),
 * 1 EQ_ELIMINATED (sparse_ldlt_eq_cond, kkt_eq_eliminated.hpp), 2 INEQ_ELIMINATED (sparse_ldlt_ineq_cond,
 * kkt_ineq_eliminated.hpp), 3 ALL_ELIMINATED (sparse_ldlt_cond, kkt_all_eliminated.hpp).
 * perm: optional fill-reducing ordering of the KKT of that mode (n + [p] + [m] entries; NULL = own AMD). */
int b200kkt_sparse_create(b200kkt_handle** out, int n, int p, int m,
                          const int* Pp, const int* Pi, const double* Px,
                          const int* ATp, const int* ATi, const double* ATx,
                          const int* GTp, const int* GTi, const double* GTx,
                          int mode, const int* perm, int device);
/* symbolic statistics of a sparse_ldlt handle (what LDLt::factorize_symbolic computes, sparse/ldlt.hpp:42-99) and the
 * ordering in use (ordering.hpp:59-125): perm[n+p+m], perm[new] = old.  Any pointer may be NULL.                    */
/* host-only symbolic phase of the sparse_ldlt backend (no GPU needed): KKT pattern (kkt_full.hpp:39-170), fill-reducing
 * ordering (ordering.hpp:59-125; perm_in != NULL uses the caller's), elimination tree / nnz(L) (ldlt.hpp:42-99) and the
 * number of etree level sets the device factorisation is scheduled by.                                              */
int b200_sparse_ldlt_symbolic(int n, int p, int m, const int* Pp, const int* Pi, const int* ATp, const int* ATi,
                              const int* GTp, const int* GTi, const int* perm_in, int* perm_out,
                              long long* nnz_kkt, long long* nnz_L, int* etree_levels, double* factor_flops);
/* the same for any KKTMode, plus the supernodal schedule's size (number of supernodes after relaxed amalgamation, rows of the
 * largest front).  perm_in / perm_out have n + [p] + [m] entries for that mode.                                      */
int b200_sparse_ldlt_symbolic_mode(int n, int p, int m, const int* Pp, const int* Pi, const int* ATp, const int* ATi,
                                   const int* GTp, const int* GTi, int mode, const int* perm_in, int* perm_out,
                                   long long* nnz_kkt, long long* nnz_L, int* etree_levels, double* factor_flops,
                                   int* n_supernodes, int* largest_front);
int b200kkt_sparse_info(b200kkt_handle* h, long long* nnz_kkt, long long* nnz_L, int* etree_levels, int* perm);

/* replaces sparse::MultistageKKT<T,I>::MultistageKKT(const Data&) (include/piqp/sparse/multistage_kkt.hpp:74-133) */
int b200kkt_multistage_create(b200kkt_handle** out, int n, int p, int m,
                              const int* Pp, const int* Pi, const double* Px,
                              const int* ATp, const int* ATi, const double* ATx,
                              const int* GTp, const int* GTi, const double* GTx, int device);

/* replaces KKTSolverBase::update_data(data, options)  (kkt_solver_base.hpp:30; dense/kkt.hpp:62-71;
 * sparse/kkt_full.hpp:212-251).  Pointers to the NEW (scaled) values, same shapes/patterns as at create;
 * only the pieces named in `options` are read.                                                       */
int b200kkt_update_data(b200kkt_handle* h, int options, const double* P, const double* AT, const double* GT);

/* replaces KKTSolverBase::update_scalings_and_factor(data, delta, x_reg, z_reg)  (kkt_solver_base.hpp:32)
 * returns 1 = factorisation succeeded, 0 = failed (non-positive / zero pivot), <0 = error.          */
int b200kkt_factor(b200kkt_handle* h, double delta, const double* x_reg, const double* z_reg);

/* replaces KKTSolverBase::solve(data, rhs_x, rhs_y, rhs_z, lhs_x, lhs_y, lhs_z)  (kkt_solver_base.hpp:34) */
int b200kkt_solve(b200kkt_handle* h, const double* rhs_x, const double* rhs_y, const double* rhs_z,
                  double* lhs_x, double* lhs_y, double* lhs_z);

/* replaces KKTSolverBase::eval_P_x(data, alpha, x, z): z = alpha P x  (kkt_solver_base.hpp:37) */
int b200kkt_eval_P_x(b200kkt_handle* h, double alpha, const double* x, double* z);
/* replaces KKTSolverBase::eval_A_xn_and_AT_xt: zn = alpha_n A xn, zt = alpha_t A^T xt  (kkt_solver_base.hpp:39) */
int b200kkt_eval_A_xn_and_AT_xt(b200kkt_handle* h, double alpha_n, double alpha_t,
                                const double* xn, const double* xt, double* zn, double* zt);
/* replaces KKTSolverBase::eval_G_xn_and_GT_xt  (kkt_solver_base.hpp:41) */
int b200kkt_eval_G_xn_and_GT_xt(b200kkt_handle* h, double alpha_n, double alpha_t,
                                const double* xn, const double* xt, double* zn, double* zt);

/* replaces KKTSolverBase::clone()  (kkt_solver_base.hpp:28): deep copy incl. factor storage */
b200kkt_handle* b200kkt_clone(const b200kkt_handle* h);
/* replaces KKTSolverBase::print_info()  (kkt_solver_base.hpp:43) */
void b200kkt_print_info(const b200kkt_handle* h);
void b200kkt_destroy(b200kkt_handle* h);

/* test/diagnostic access: copy the assembled lower-triangular n x n KKT matrix (dense backend only,
 * the twin of dense::KKT::internal_kkt_mat(), dense/kkt.hpp:134-137) and its Cholesky factor.        */
int b200kkt_dense_get_kkt(b200kkt_handle* h, double* kkt_lower, double* chol_lower);

/* ------------------------------------------------------------------------------------------------
 * (2) batched QP solver with device-resident interior-point loop
 * ------------------------------------------------------------------------------------------------ */

/* layout-identical to piqp_settings (interfaces/c/include/piqp_typedef.h:75-104) */
typedef struct {
    double rho_init;
    double delta_init;
    double eps_abs;
    double eps_rel;
    int check_duality_gap;
    double eps_duality_gap_abs;
    double eps_duality_gap_rel;
    double infeasibility_threshold;
    double reg_lower_limit;
    double reg_finetune_lower_limit;
    int reg_finetune_primal_update_threshold;
    int reg_finetune_dual_update_threshold;
    int max_iter;
    int max_factor_retires;
    int preconditioner_scale_cost;
    int preconditioner_reuse_on_update;
    int preconditioner_iter;
    double tau;
    int kkt_solver; /* piqp_kkt_solver: 0 dense_cholesky, 1 sparse_ldlt, ..., 5 sparse_multistage */
    int iterative_refinement_always_enabled;
    double iterative_refinement_eps_abs;
    double iterative_refinement_eps_rel;
    int iterative_refinement_max_iter;
    double iterative_refinement_min_improvement_rate;
    double iterative_refinement_static_regularization_eps;
    double iterative_refinement_static_regularization_rel;
    int verbose;
    int compute_timings;
} b200qp_settings;

/* layout-identical to piqp_info (interfaces/c/include/piqp_typedef.h:116-159) */
typedef struct {
    int status; /* piqp_status: 1 solved, -1 max iter, -2 primal inf., -3 dual inf., -8 numerics, -9 unsolved, -10 invalid settings */
    int iter;
    double rho, delta, mu, sigma, primal_step, dual_step;
    double primal_res, primal_res_rel, dual_res, dual_res_rel;
    double primal_res_reg, primal_res_reg_rel, dual_res_reg, dual_res_reg_rel;
    double primal_prox_inf, dual_prox_inf;
    double prev_primal_res, prev_dual_res;
    double primal_obj, dual_obj, duality_gap, duality_gap_rel;
    int factor_retires;
    double reg_limit;
    int no_primal_update;
    int no_dual_update;
    double setup_time, update_time, solve_time, kkt_factor_time, kkt_solve_time, run_time;
} b200qp_info;

/* aggregate counters of the last b200qp_solve (ours; used for the GFLOP/s accounting) */
typedef struct {
    long long factor_calls;     /* sum over instances of update_scalings_and_factor calls        */
    long long kkt_solve_calls;  /* sum over instances of KKTSystem::solve calls                   */
    long long backend_solves;   /* sum over instances of backend solve calls (incl. refinement)   */
    long long ip_iterations;    /* sum over instances of info.iter                                */
    int lockstep_iterations;    /* number of batched iterations the host loop ran                 */
    double factor_ms;           /* device time in the factor bucket (CUDA events)                 */
    double solve_ms;            /* device time in the KKT-solve bucket                             */
    double total_ms;            /* device time of the whole solve                                  */
    unsigned long long kernel_launches;
    /* per-kernel-class device times of the last solve (only when profiling was enabled with
     * b200qp_set_profiling): KKT assembly contraction, Cholesky, backend solve */
    double assemble_ms, cholesky_ms, backend_solve_ms;
    long long assemble_launches, cholesky_calls, backend_solve_launch_groups;
} b200qp_stats;

typedef struct b200qp_handle b200qp_handle;

/* piqp_set_default_settings_dense / _sparse (piqp.h:24-25) */
void b200qp_set_default_settings_dense(b200qp_settings* s);
void b200qp_set_default_settings_sparse(b200qp_settings* s);

/* Batched twin of piqp_setup_dense (piqp.h:27): `batch` independent QPs of identical shape (n, p, m).
 * All arrays are instance-major: P[batch][n][n] (row-major, upper triangle used), c[batch][n],
 * A[batch][p][n], b[batch][p], G[batch][m][n], h_l/h_u[batch][m], x_l/x_u[batch][n]; A,b,G,h_l,h_u,x_l,x_u
 * may be NULL like in the reference.  `on_device` != 0: the pointers are device pointers on `device`
 * (inputs already resident in HBM); otherwise host pointers (pinned or pageable) copied inside the call. */
int b200qp_setup_dense(b200qp_handle** out, int batch, int n, int p, int m,
                       const double* P, const double* c, const double* A, const double* b,
                       const double* G, const double* h_l, const double* h_u,
                       const double* x_l, const double* x_u,
                       const b200qp_settings* settings, int device, int on_device);

/* Batched twin of piqp_update_dense (piqp.h:31): NULL = keep. */
int b200qp_update_dense(b200qp_handle* h, const double* P, const double* c, const double* A, const double* b,
                        const double* G, const double* h_l, const double* h_u,
                        const double* x_l, const double* x_u, int on_device);

/* Batched twin of piqp_setup_sparse (piqp.h:28, piqp_data_sparse piqp_typedef.h:56-68): all instances share the CSC
 * patterns of P (n x n, upper triangle used), A (p x n) and G (m x n) -- int32 column pointers / row indices on the host --
 * and differ in the value arrays Px[batch][nnz(P)], Ax[batch][nnz(A)], Gx[batch][nnz(G)] and in the vectors.
 * settings->kkt_solver selects the backend (settings.hpp:18-26): 1 sparse_ldlt, 2 sparse_ldlt_eq_cond,
 * 3 sparse_ldlt_ineq_cond, 4 sparse_ldlt_cond (supernodal multifrontal LDL^T on the KKT of that mode),
 * 5 sparse_multistage (block-tridiagonal-arrow Cholesky).                                                           */
int b200qp_setup_sparse(b200qp_handle** out, int batch, int n, int p, int m,
                        const int* Pp, const int* Pi, const double* Px, const double* c,
                        const int* Ap, const int* Ai, const double* Ax, const double* b,
                        const int* Gp, const int* Gi, const double* Gx, const double* h_l, const double* h_u,
                        const double* x_l, const double* x_u, const b200qp_settings* settings, int device, int on_device);
/* The same with a caller-supplied fill-reducing ordering of the KKT matrix of the selected mode (kkt_perm[new] = old,
 * n + [p] + [m] entries, sparse/ordering.hpp:59-125; NULL = own AMD).  In the multi-GPU mode rank 0 runs the symbolic analysis
 * once and the permutation rides the setup broadcast (b200qp_get_sparse_perm on rank 0 -> NCCL -> kkt_perm elsewhere). */
int b200qp_setup_sparse_ex(b200qp_handle** out, int batch, int n, int p, int m,
                           const int* Pp, const int* Pi, const double* Px, const double* c,
                           const int* Ap, const int* Ai, const double* Ax, const double* b,
                           const int* Gp, const int* Gi, const double* Gx, const double* h_l, const double* h_u,
                           const double* x_l, const double* x_u, const b200qp_settings* settings, int device, int on_device, const int* kkt_perm);
/* ordering in use by a sparse_ldlt-family handle: copies min(cap, n_kkt) entries, returns n_kkt */
int b200qp_get_sparse_perm(b200qp_handle* h, int* perm, int cap);
/* Batched twin of piqp_update_sparse (piqp.h:36): same patterns, new values; NULL = keep. */
int b200qp_update_sparse(b200qp_handle* h, const double* Px, const double* c, const double* Ax, const double* b, const double* Gx,
                         const double* h_l, const double* h_u, const double* x_l, const double* x_u, int on_device);
/* detected multistage structure as (start, diag_size, off_diag_size) triples, last = arrow block
 * (what MultistageKKT::print_info prints, multistage_kkt.hpp:385-392); returns the number of blocks */
int b200qp_multistage_blocks(b200qp_handle* h, int* out, int cap);
int b200kkt_multistage_blocks(b200kkt_handle* h, int* out, int cap);

/* piqp_update_settings (piqp.h:30) */
int b200qp_update_settings(b200qp_handle* h, const b200qp_settings* settings);

/* Batched twin of piqp_solve (piqp.h:42): runs all instances to termination. Returns 0 or an error;
 * per-instance piqp_status values are in the infos.                                               */
int b200qp_solve(b200qp_handle* h);

/* Results, instance-major, in the reference's public layout (piqp_result, piqp_typedef.h:161-175):
 * x[batch][n] y[batch][p] z_l,z_u,s_l,s_u[batch][m] z_bl,z_bu,s_bl,s_bu[batch][n]; any pointer may be NULL.
 * `on_device`: destination pointers are device pointers.                                          */
int b200qp_get_result(b200qp_handle* h, double* x, double* y, double* z_l, double* z_u, double* z_bl, double* z_bu,
                      double* s_l, double* s_u, double* s_bl, double* s_bu, int on_device);
int b200qp_get_info(b200qp_handle* h, b200qp_info* infos /* [batch] */);
int b200qp_get_stats(b200qp_handle* h, b200qp_stats* stats);
/* per-iteration trace of instance `b`: rows of (rho, delta, mu, primal_step, dual_step, primal_res, dual_res,
 * primal_obj, dual_obj, duality_gap); returns number of rows (needs settings.verbose >= 2 at setup). */
int b200qp_get_trace(b200qp_handle* h, int b, double* rows, int max_rows);
/* algorithmic work of ONE backend call for ONE instance (SURVEY.md 8d formulas; for multistage the reference's own
 * cost model, multistage_kkt.hpp:397-418): flops and bytes of update_scalings_and_factor and of the backend solve */
int b200qp_get_work(b200qp_handle* h, double* factor_flops, double* factor_bytes, double* solve_flops, double* solve_bytes);
/* enable / disable CUDA-event timing of the backend's kernel classes (adds two event records per call) */
int b200qp_set_profiling(b200qp_handle* h, int enable);
void b200qp_cleanup(b200qp_handle* h);

/* Bench hooks: run `reps` factor calls (assemble + factorise) and `nsolve` backend solves per factor on the
 * CURRENT scalings of all instances and return device milliseconds per bucket (CUDA events on the
 * library's stream).  Used by bench.py for the kernel-level roofline numbers.                      */
int b200qp_bench_factor_solve(b200qp_handle* h, int reps, int nsolve, double* factor_ms, double* solve_ms);

#ifdef __cplusplus
}
#endif
#endif /* PIQP_B200_H */
