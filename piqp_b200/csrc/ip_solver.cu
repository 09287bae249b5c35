// piqp_b200/csrc/ip_solver.cu -- batched, device-resident interior-point driver (see ip_solver.hpp).
// One CTA per QP instance for every phase kernel; O(n+m) vector work + block reductions.  The arithmetic
// order of the element-wise formulas follows the reference line by line (citations inline) so that
// iteration counts match the CPU solver; only reductions (dot / min / max) use a fixed tree instead of a
// sequential loop.
#include "ip_solver.hpp"
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <mutex>

namespace b200 {

constexpr int IPT = 256;
#define PB(ptr, len) ((ptr) + (size_t)b * (size_t)(len))
#define FOR_T(i, len) for (int i = threadIdx.x; i < (len); i += blockDim.x)

enum { ST_SOLVED = 1, ST_MAX_ITER = -1, ST_PRIMAL_INF = -2, ST_DUAL_INF = -3, ST_NUMERICS = -8, ST_UNSOLVED = -9, ST_INVALID = -10 };

struct InstPtr {   // per-instance views
    const double *c, *bv, *h_l, *h_u, *x_l, *x_u, *xbs;
    const int *hhl, *hhu, *hxl, *hxu;
    const double *pd, *pd_inv, *pdb, *pdb_inv;
    double pc, pc_inv;
};
__device__ __forceinline__ InstPtr inst(const IpDev& d, int b) {
    InstPtr q;
    const int N = d.n + d.p + d.m;
    q.c = PB(d.c, d.n); q.bv = PB(d.b, d.p); q.h_l = PB(d.h_l, d.m); q.h_u = PB(d.h_u, d.m);
    q.x_l = PB(d.x_l, d.n); q.x_u = PB(d.x_u, d.n); q.xbs = PB(d.xbs, d.n);
    q.hhl = PB(d.has_hl, d.m); q.hhu = PB(d.has_hu, d.m); q.hxl = PB(d.has_xl, d.n); q.hxu = PB(d.has_xu, d.n);
    q.pd = PB(d.pd, N); q.pd_inv = PB(d.pd_inv, N); q.pdb = PB(d.pdb, d.n); q.pdb_inv = PB(d.pdb_inv, d.n);
    q.pc = d.pc[b]; q.pc_inv = d.pc_inv[b];
    return q;
}
struct VarsB { double *x, *y, *z_l, *z_u, *z_bl, *z_bu, *s_l, *s_u, *s_bl, *s_bu; };
__device__ __forceinline__ VarsB vb(const Vars& v, const IpDev& d, int b) {
    VarsB r;
    r.x = PB(v.x, d.n); r.y = PB(v.y, d.p); r.z_l = PB(v.z_l, d.m); r.z_u = PB(v.z_u, d.m); r.z_bl = PB(v.z_bl, d.n); r.z_bu = PB(v.z_bu, d.n);
    r.s_l = PB(v.s_l, d.m); r.s_u = PB(v.s_u, d.m); r.s_bl = PB(v.s_bl, d.n); r.s_bu = PB(v.s_bu, d.n);
    return r;
}

// preconditioner element maps (dense/preconditioner.hpp:253-421)
#define US_PRIMAL(v, i) ((v) * q.pd[i])
#define US_DUAL_EQ(v, i) ((v) * q.pc_inv * q.pd[d.n + (i)])
#define US_DUAL_INEQ(v, i) ((v) * q.pc_inv * q.pd[d.n + d.p + (i)])
#define US_DUAL_B(v, i) ((v) * q.pc_inv * q.pdb[i])
#define US_SLACK_INEQ(v, i) ((v) * q.pd_inv[d.n + d.p + (i)])
#define US_SLACK_B(v, i) ((v) * q.pdb_inv[i])
#define US_PRES_EQ(v, i) ((v) * q.pd_inv[d.n + (i)])
#define US_PRES_INEQ(v, i) ((v) * q.pd_inv[d.n + d.p + (i)])
#define US_PRES_B(v, i) ((v) * q.pdb_inv[i])
#define US_DRES(v, i) ((v) * q.pc_inv * q.pd_inv[i])

__shared__ double g_red[32 * 12];   // block_reduce scratch: up to 12 values per call

// ---------------------------------------------------------------------------------------------------
__global__ void k_counts(IpDev d) {   // n_fin = n_h_l + n_h_u + n_x_l + n_x_u ; has_ineq = m + n_x_l + n_x_u > 0
    const int b = blockIdx.x;
    InstPtr q = inst(d, b);
    double v[2] = {0.0, 0.0};
    FOR_T(i, d.m) v[0] += (q.hhl[i] ? 1.0 : 0.0) + (q.hhu[i] ? 1.0 : 0.0);
    FOR_T(i, d.n) v[1] += (q.hxl[i] ? 1.0 : 0.0) + (q.hxu[i] ? 1.0 : 0.0);
    const int op[2] = {RED_SUM, RED_SUM};
    block_reduce<2>(v, op, g_red);
    if (threadIdx.x == 0) { d.sc[b].n_fin = v[0] + v[1]; d.sc[b].has_ineq = (d.m + (int)v[1]) > 0; }
}

__global__ void k_init(IpDev d) {   // solver.hpp:398-439
    const int b = blockIdx.x;
    InstPtr q = inst(d, b);
    VarsB it = vb(d.it, d, b);
    FOR_T(i, d.m) { const double l = q.hhl[i] ? 1.0 : 0.0, u = q.hhu[i] ? 1.0 : 0.0; it.s_l[i] = l; it.z_l[i] = l; it.s_u[i] = u; it.z_u[i] = u; }
    FOR_T(i, d.n) { const double l = q.hxl[i] ? 1.0 : 0.0, u = q.hxu[i] ? 1.0 : 0.0; it.s_bl[i] = l; it.z_bl[i] = l; it.s_bu[i] = u; it.z_bu[i] = u; it.x[i] = 0.0; }
    FOR_T(i, d.p) it.y[i] = 0.0;
    if (threadIdx.x == 0) {
        IpScalars& s = d.sc[b];
        s.status = ST_UNSOLVED; s.iter = 0; s.reg_limit = d.st.reg_lower_limit; s.factor_retires = 0;
        s.no_primal_update = 0; s.no_dual_update = 0; s.mu = 0; s.sigma = 0; s.primal_step = 0; s.dual_step = 0;
        s.rho = d.st.rho_init; s.delta = d.st.delta_init;
        s.primal_res = 0; s.dual_res = 0; s.primal_res_rel = 0; s.dual_res_rel = 0; s.prev_primal_res = 0; s.prev_dual_res = 0;
        s.primal_res_reg = 0; s.primal_res_reg_rel = 0; s.dual_res_reg = 0; s.dual_res_reg_rel = 0; s.primal_prox_inf = 0; s.dual_prox_inf = 0;
        s.primal_obj = 0; s.dual_obj = 0; s.duality_gap = 0; s.duality_gap_rel = 0; s.mu_rate = 0;
        s.ir_on = d.st.iterative_refinement_always_enabled ? 1 : 0;
        s.active = 1; s.need_factor = 1; s.reg_changed = 0; s.use_ir = 0; s.ir_continue = 0;
        s.n_factor = 0; s.n_solve = 0; s.n_backend_solve = 0;
        d.act[b] = 1; d.act2[b] = 0; d.need_factor[b] = 1; d.need_factor[d.batch + b] = s.ir_on; d.ok[b] = 0; d.ir_mask[b] = 0;
    }
}

// KKTSystem::update_scalings_and_factor, vector part (kkt_system.hpp:143-211)
__global__ void k_prepare_factor(IpDev d) {
    const int b = blockIdx.x;
    if (!d.need_factor[b]) return;
    InstPtr q = inst(d, b);
    VarsB it = vb(d.it, d, b);
    IpScalars& s = d.sc[b];
    const double rho = s.rho, delta = s.delta;
    double *ksl = PB(d.k_s_l, d.m), *ksu = PB(d.k_s_u, d.m), *kzl = PB(d.k_zl_inv, d.m), *kzu = PB(d.k_zu_inv, d.m);
    double *ksbl = PB(d.k_s_bl, d.n), *ksbu = PB(d.k_s_bu, d.n), *kzbl = PB(d.k_zbl_inv, d.n), *kzbu = PB(d.k_zbu_inv, d.n);
    double *xr = PB(d.x_reg, d.n), *zr = PB(d.z_reg, d.m), *zri = PB(d.z_reg_ir, d.m);
    const double* pdg = PB(d.P_diag, d.n);
    double v[2] = {0.0, 0.0};
    FOR_T(i, d.n) {
        double xv = rho;
        const int fl = q.hxl[i], fu = q.hxu[i];
        const double sl = it.s_bl[i], zl = it.z_bl[i], su = it.s_bu[i], zu = it.z_bu[i], xb = q.xbs[i], pdi = pdg[i];
        if (fl) { const double kz = 1.0 / zl; ksbl[i] = sl; kzbl[i] = kz; xv += xb * xb / (kz * sl + delta); }
        if (fu) { const double kz = 1.0 / zu; ksbu[i] = su; kzbu[i] = kz; xv += xb * xb / (kz * su + delta); }
        xr[i] = xv;
        v[0] = fmax(v[0], fabs(pdi + xv));
    }
    FOR_T(i, d.m) {
        ksl[i] = it.s_l[i]; ksu[i] = it.s_u[i]; kzl[i] = 1.0 / it.z_l[i]; kzu[i] = 1.0 / it.z_u[i];
        double zv = 0.0;
        if (q.hhl[i]) zv += 1.0 / (kzl[i] * ksl[i] + delta);
        if (q.hhu[i]) zv += 1.0 / (kzu[i] * ksu[i] + delta);
        zv = 1.0 / zv;
        zr[i] = zv; zri[i] = zv;
        v[1] = fmax(v[1], fabs(zv));
    }
    double delta_reg = delta;
    if (s.ir_on) {   // block-uniform
        const int op[2] = {RED_MAX, RED_MAX};
        block_reduce<2>(v, op, g_red);
        const double max_diag = fmax(v[0], v[1]);
        const double reg = d.st.iterative_refinement_static_regularization_eps + d.st.iterative_refinement_static_regularization_rel * max_diag;
        delta_reg += reg;
        FOR_T(i, d.n) xr[i] += reg;
        FOR_T(i, d.m) zri[i] += reg;
    }
    if (threadIdx.x == 0) { s.use_ir = s.ir_on; s.kkt_rho = rho; s.kkt_delta = delta; d.delta_reg[b] = delta_reg; }
}

// factor retry ladder (solver.hpp:446-465, 688-708)
__global__ void k_after_factor(IpDev d) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= d.batch || !d.need_factor[b]) return;
    IpScalars& s = d.sc[b];
    s.n_factor++;
    if (d.ok[b]) { s.need_factor = 0; s.factor_retires = 0; }
    else if (!s.ir_on) { s.ir_on = 1; }
    else if (s.factor_retires < d.st.max_factor_retires) {
        s.delta *= 100; s.rho *= 100; s.factor_retires++;
        s.reg_limit = fmin(10 * s.reg_limit, d.st.eps_abs);
        s.reg_changed = 1;
    } else { s.status = ST_NUMERICS; s.active = 0; s.need_factor = 0; d.act[b] = 0; d.act2[b] = 0; }
    d.need_factor[b] = s.need_factor;
    d.need_factor[d.batch + b] = s.ir_on;   // second half: refinement flags, read back together
}

__global__ void k_initial_rhs(IpDev d) {   // solver.hpp:473-482
    const int b = blockIdx.x;
    if (!d.act[b]) return;
    InstPtr q = inst(d, b);
    VarsB r = vb(d.r, d, b);
    FOR_T(i, d.n) { r.x[i] = -q.c[i]; r.z_bl[i] = -q.x_l[i]; r.z_bu[i] = q.x_u[i]; r.s_bl[i] = 0; r.s_bu[i] = 0; }
    FOR_T(i, d.p) r.y[i] = q.bv[i];
    FOR_T(i, d.m) { r.z_l[i] = -q.h_l[i]; r.z_u[i] = q.h_u[i]; r.s_l[i] = 0; r.s_u[i] = 0; }
}

// KKTSystem::solve prologue (kkt_system.hpp:219-252)
__global__ void k_solve_pre(IpDev d, Vars rhsv, const int* mask) {
    const int b = blockIdx.x;
    if (!mask[b]) return;
    InstPtr q = inst(d, b);
    VarsB rhs = vb(rhsv, d, b);
    const double delta = d.sc[b].kkt_delta;
    const double *ksl = PB(d.k_s_l, d.m), *ksu = PB(d.k_s_u, d.m), *kzl = PB(d.k_zl_inv, d.m), *kzu = PB(d.k_zu_inv, d.m);
    const double *ksbl = PB(d.k_s_bl, d.n), *ksbu = PB(d.k_s_bu, d.n), *kzbl = PB(d.k_zbl_inv, d.n), *kzbu = PB(d.k_zbu_inv, d.n);
    const double* zr = PB(d.z_reg, d.m);
    double *rzb = PB(d.rhs_z_bar, d.m), *rxb = PB(d.rhs_x_bar, d.n);
    FOR_T(i, d.m) {
        double v = 0.0;
        if (q.hhl[i]) v -= 1.0 / (kzl[i] * ksl[i] + delta) * (rhs.z_l[i] - kzl[i] * rhs.s_l[i]);
        if (q.hhu[i]) v += 1.0 / (kzu[i] * ksu[i] + delta) * (rhs.z_u[i] - kzu[i] * rhs.s_u[i]);
        rzb[i] = v * zr[i];
    }
    FOR_T(i, d.n) {
        double v = rhs.x[i];
        const int fl = q.hxl[i], fu = q.hxu[i];
        const double xb = q.xbs[i], rzl = rhs.z_bl[i], rsl = rhs.s_bl[i], kzl_ = kzbl[i], ksl_ = ksbl[i], rzu = rhs.z_bu[i], rsu = rhs.s_bu[i], kzu_ = kzbu[i], ksu_ = ksbu[i];
        if (fl) v -= xb * (rzl - kzl_ * rsl) / (ksl_ * kzl_ + delta);
        if (fu) v += xb * (rzu - kzu_ * rsu) / (ksu_ * kzu_ + delta);
        rxb[i] = v;
    }
    if (threadIdx.x == 0) { d.sc[b].n_solve++; d.sc[b].n_backend_solve++; }
}

// KKTSystem::solve epilogue: dual recovery (kkt_system.hpp:310-366)
__global__ void k_solve_post(IpDev d, Vars rhsv, Vars lhsv, const int* mask) {
    const int b = blockIdx.x;
    if (!mask[b]) return;
    InstPtr q = inst(d, b);
    VarsB rhs = vb(rhsv, d, b), lhs = vb(lhsv, d, b);
    const double delta = d.sc[b].kkt_delta;
    const double *ksl = PB(d.k_s_l, d.m), *ksu = PB(d.k_s_u, d.m), *kzl = PB(d.k_zl_inv, d.m), *kzu = PB(d.k_zu_inv, d.m);
    const double *ksbl = PB(d.k_s_bl, d.n), *ksbu = PB(d.k_s_bu, d.n), *kzbl = PB(d.k_zbl_inv, d.n), *kzbu = PB(d.k_zbu_inv, d.n);
    const double *zr = PB(d.z_reg, d.m), *lz = PB(d.lhs_z, d.m);
    FOR_T(i, d.m) {
        const bool hl = q.hhl[i], hu = q.hhu[i];
        if (hl && hu) {
            const double rzl = rhs.z_l[i] - kzl[i] * rhs.s_l[i];
            const double Wl = 1.0 / (kzl[i] * ksl[i] + delta);
            const double rzu = rhs.z_u[i] - kzu[i] * rhs.s_u[i];
            const double Wu = 1.0 / (kzu[i] * ksu[i] + delta);
            const double rs = Wl * Wu * (rzl + rzu);
            const double zl = -zr[i] * (rs + Wl * lz[i]);
            const double zu = -zr[i] * (rs - Wu * lz[i]);
            lhs.z_l[i] = zl; lhs.z_u[i] = zu;
            lhs.s_l[i] = kzl[i] * (rhs.s_l[i] - ksl[i] * zl);
            lhs.s_u[i] = kzu[i] * (rhs.s_u[i] - ksu[i] * zu);
        } else if (hl) {
            const double zl = -lz[i];
            lhs.z_l[i] = zl; lhs.z_u[i] = 0.0;
            lhs.s_l[i] = kzl[i] * (rhs.s_l[i] - ksl[i] * zl); lhs.s_u[i] = 0.0;
        } else if (hu) {
            const double zu = lz[i];
            lhs.z_l[i] = 0.0; lhs.z_u[i] = zu;
            lhs.s_l[i] = 0.0; lhs.s_u[i] = kzu[i] * (rhs.s_u[i] - ksu[i] * zu);
        }
    }
    FOR_T(i, d.n) {
        const int fl = q.hxl[i], fu = q.hxu[i];
        const double xb = q.xbs[i], lx = lhs.x[i], rzl = rhs.z_bl[i], rsl = rhs.s_bl[i], kzl_ = kzbl[i], ksl_ = ksbl[i], rzu = rhs.z_bu[i], rsu = rhs.s_bu[i], kzu_ = kzbu[i], ksu_ = ksbu[i];
        if (fl) {
            const double z = (-xb * lx - rzl + kzl_ * rsl) / (ksl_ * kzl_ + delta);
            lhs.z_bl[i] = z; lhs.s_bl[i] = kzl_ * (rsl - ksl_ * z);
        }
        if (fu) {
            const double z = (xb * lx - rzu + kzu_ * rsu) / (ksu_ * kzu_ + delta);
            lhs.z_bu[i] = z; lhs.s_bu[i] = kzu_ * (rsu - ksu_ * z);
        }
    }
}

// ---- iterative refinement (kkt_system.hpp:256-301, 507-536) ------------------------------------------
// err = rhs - K3x3 * l, with P*l in err_x, (A l_x) in err_y, (G l_x) in err_z, A^T l_y in work_x, G^T l_z in work_x2.
// stage 0: first error of this solve; stage 1: error of the candidate ref_* and accept / continue decision.
__global__ void k_ir_err(IpDev d, Vars rhsv, const double* lx_, const double* ly_, const double* lz_, int stage, int it_idx, const int* mask) {
    const int b = blockIdx.x;
    if (!mask[b]) return;
    VarsB rhs = vb(rhsv, d, b);
    IpScalars& s = d.sc[b];
    const double *lx = PB(lx_, d.n), *ly = PB(ly_, d.p), *lz = PB(lz_, d.m);
    const double *xr = PB(d.x_reg, d.n), *zr = PB(d.z_reg, d.m), *wx = PB(d.work_x, d.n), *wx2 = PB(d.work_x2, d.n);
    const double *rxb = PB(d.rhs_x_bar, d.n), *rzb = PB(d.rhs_z_bar, d.m);
    double *ex = PB(d.err_x, d.n), *ey = PB(d.err_y, d.p), *ez = PB(d.err_z, d.m);
    const double delta = s.kkt_delta;
    double v[2] = {0.0, 0.0};   // err norm, rhs norm
    FOR_T(i, d.n) {
        double e = ex[i]; e += xr[i] * lx[i]; e += wx[i]; e += wx2[i];
        e = rxb[i] - e; ex[i] = e;
        v[0] = fmax(v[0], fabs(e)); v[1] = fmax(v[1], fabs(rxb[i]));
    }
    FOR_T(i, d.p) {
        double e = ey[i]; e -= delta * ly[i];
        e = rhs.y[i] - e; ey[i] = e;
        v[0] = fmax(v[0], fabs(e)); v[1] = fmax(v[1], fabs(rhs.y[i]));
    }
    FOR_T(i, d.m) {
        double e = ez[i]; e -= zr[i] * lz[i];
        e = rzb[i] - e; ez[i] = e;
        v[0] = fmax(v[0], fabs(e)); v[1] = fmax(v[1], fabs(rzb[i]));
    }
    // NaN-aware max: fmax drops NaNs, so test finiteness separately
    double bad = 0.0;
    FOR_T(i, d.n) if (!isfinite(ex[i])) bad = 1.0;
    FOR_T(i, d.p) if (!isfinite(ey[i])) bad = 1.0;
    FOR_T(i, d.m) if (!isfinite(ez[i])) bad = 1.0;
    double w[3] = {v[0], v[1], bad};
    const int op[3] = {RED_MAX, RED_MAX, RED_MAX};
    block_reduce<3>(w, op, g_red);
    const double err = w[2] > 0.0 ? INFINITY : w[0];
    const double tol_abs = d.st.iterative_refinement_eps_abs, tol_rel = d.st.iterative_refinement_eps_rel;
    int cont = 0, accept = 0;
    if (stage == 0) {
        const double rhs_norm = w[1];
        cont = isfinite(err) && !(err <= tol_abs + tol_rel * rhs_norm) && (d.st.iterative_refinement_max_iter > 0);
        if (threadIdx.x == 0) { s.rhs_norm = rhs_norm; s.refine_err = err; }
    } else {
        const double prev = s.refine_err;
        if (isfinite(err)) {
            const double rate = prev / err;
            if (rate < d.st.iterative_refinement_min_improvement_rate) { accept = rate > 1.0; cont = 0; }
            else {
                accept = 1;
                cont = !(err <= tol_abs + tol_rel * s.rhs_norm) && (it_idx + 1 < d.st.iterative_refinement_max_iter);
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) s.refine_err = err;
    }
    if (threadIdx.x == 0) { s.ir_continue = cont; d.ir_mask[b] = cont; d.ok[b] = accept; if (cont) s.n_backend_solve++; }
}
// ref += lhs (kkt_system.hpp:277-280)
__global__ void k_ir_accum(IpDev d, const double* lx_, const double* ly_, const double* lz_, const int* mask) {
    const int b = blockIdx.x;
    if (!mask[b]) return;
    double *rx = PB(d.ref_x, d.n), *ry = PB(d.ref_y, d.p), *rz = PB(d.ref_z, d.m);
    FOR_T(i, d.n) rx[i] += PB(lx_, d.n)[i];
    FOR_T(i, d.p) ry[i] += PB(ly_, d.p)[i];
    FOR_T(i, d.m) rz[i] += PB(lz_, d.m)[i];
}
// lhs <- ref for instances whose candidate was accepted (flag in d.ok); `was` = instances that ran this round
__global__ void k_ir_accept(IpDev d, double* lx_, double* ly_, double* lz_, const int* was) {
    const int b = blockIdx.x;
    if (!was[b] || !d.ok[b]) return;
    const double *rx = PB(d.ref_x, d.n), *ry = PB(d.ref_y, d.p), *rz = PB(d.ref_z, d.m);
    FOR_T(i, d.n) PB(lx_, d.n)[i] = rx[i];
    FOR_T(i, d.p) PB(ly_, d.p)[i] = ry[i];
    FOR_T(i, d.m) PB(lz_, d.m)[i] = rz[i];
}
__global__ void k_mask_and_flag(const int* mask, IpDev d, int* out) {   // out[b] = mask[b] && use_ir
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < d.batch) out[b] = mask[b] && d.sc[b].use_ir;
}
__global__ void k_copy_int(const int* src, int* dst, int n) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < n) dst[b] = src[b];
}

// ---- shared pieces ------------------------------------------------------------------------------------
__device__ double calc_mu(const IpDev& d, const InstPtr& q, const VarsB& it, double n_fin) {   // solver.hpp:884-891
    double v[4] = {0, 0, 0, 0};
    FOR_T(i, d.m) { v[0] += it.s_l[i] * it.z_l[i]; v[1] += it.s_u[i] * it.z_u[i]; }
    FOR_T(i, d.n) {      // loads first, flags second: a load behind a data-dependent branch costs one L2 round trip per level (see the note at FOR_T)
        const int fl = q.hxl[i], fu = q.hxu[i];
        const double sl = it.s_bl[i], zl = it.z_bl[i], su = it.s_bu[i], zu = it.z_bu[i];
        if (fl) v[2] += sl * zl;
        if (fu) v[3] += su * zu;
    }
    const int op[4] = {RED_SUM, RED_SUM, RED_SUM, RED_SUM};
    block_reduce<4>(v, op, g_red);
    return (v[0] + v[1] + v[2] + v[3]) / n_fin;
}

__device__ void calc_step(const IpDev& d, const InstPtr& q, const VarsB& it, const VarsB& st, double& as, double& az) {   // solver.hpp:893-958
    double v[2] = {1.0, 1.0};
    FOR_T(i, d.m) {
        const double dsl = st.s_l[i], dsu = st.s_u[i], dzl = st.z_l[i], dzu = st.z_u[i];
        const double sl = it.s_l[i], su = it.s_u[i], zl = it.z_l[i], zu = it.z_u[i];
        if (dsl < 0) v[0] = fmin(v[0], -sl / dsl);
        if (dsu < 0) v[0] = fmin(v[0], -su / dsu);
        if (dzl < 0) v[1] = fmin(v[1], -zl / dzl);
        if (dzu < 0) v[1] = fmin(v[1], -zu / dzu);
    }
    FOR_T(i, d.n) {
        const int fl = q.hxl[i], fu = q.hxu[i];
        const double dsl = st.s_bl[i], dsu = st.s_bu[i], dzl = st.z_bl[i], dzu = st.z_bu[i];
        const double sl = it.s_bl[i], su = it.s_bu[i], zl = it.z_bl[i], zu = it.z_bu[i];
        if (fl) { if (dsl < 0) v[0] = fmin(v[0], -sl / dsl); if (dzl < 0) v[1] = fmin(v[1], -zl / dzl); }
        if (fu) { if (dsu < 0) v[0] = fmin(v[0], -su / dsu); if (dzu < 0) v[1] = fmin(v[1], -zu / dzu); }
    }
    const int op[2] = {RED_MIN, RED_MIN};
    block_reduce<2>(v, op, g_red);
    as = v[0]; az = v[1];
}

// update_residuals_r (solver.hpp:1107-1128) with primal_res_r / dual_res_r / prox infs (:1148-1203).
// Results are returned to ALL threads (block-uniform) through `o`.
struct ResR { double primal_res_reg, primal_res_reg_rel, dual_res_reg, dual_res_reg_rel, primal_prox_inf, dual_prox_inf; };
__device__ ResR residuals_r(const IpDev& d, int b, const InstPtr& q, double rho, double delta, double primal_res, double primal_res_rel, double dual_res, double dual_res_rel) {
    VarsB it = vb(d.it, d, b), r = vb(d.r, d, b), rnr = vb(d.rnr, d, b), px = vb(d.prox, d, b);
    double v[4] = {0, 0, 0, 0};   // primal_res_reg, dual_res_reg, primal_prox_inf, dual_prox_inf
    FOR_T(i, d.n) {
        const double rx = rnr.x[i] - rho * (it.x[i] - px.x[i]); r.x[i] = rx;
        v[1] = fmax(v[1], fabs(US_DRES(rx, i)));
        v[3] = fmax(v[3], fabs(US_PRIMAL(it.x[i] - px.x[i], i)));
    }
    FOR_T(i, d.p) {
        const double ry = rnr.y[i] - delta * (px.y[i] - it.y[i]); r.y[i] = ry;
        v[0] = fmax(v[0], fabs(US_PRES_EQ(ry, i)));
        v[2] = fmax(v[2], fabs(US_DUAL_EQ(px.y[i] - it.y[i], i)));
    }
    FOR_T(i, d.m) {
        const double rl = rnr.z_l[i] - delta * (px.z_l[i] - it.z_l[i]); r.z_l[i] = rl;
        const double ru = rnr.z_u[i] - delta * (px.z_u[i] - it.z_u[i]); r.z_u[i] = ru;
        v[0] = fmax(v[0], fmax(fabs(US_PRES_INEQ(rl, i)), fabs(US_PRES_INEQ(ru, i))));
        v[2] = fmax(v[2], fmax(fabs(US_DUAL_INEQ(px.z_l[i] - it.z_l[i], i)), fabs(US_DUAL_INEQ(px.z_u[i] - it.z_u[i], i))));
    }
    FOR_T(i, d.n) {   // signed (no abs) for the box terms, as in the reference
        const int fl = q.hxl[i], fu = q.hxu[i];
        const double nl = rnr.z_bl[i], pl = px.z_bl[i], zl = it.z_bl[i], nu = rnr.z_bu[i], pu = px.z_bu[i], zu = it.z_bu[i];
        const double pbi = q.pdb_inv[i], pb = q.pdb[i];
        if (fl) { const double rb = nl - delta * (pl - zl); r.z_bl[i] = rb;
            v[0] = fmax(v[0], rb * pbi); v[2] = fmax(v[2], (pl - zl) * q.pc_inv * pb); }
        if (fu) { const double rb = nu - delta * (pu - zu); r.z_bu[i] = rb;
            v[0] = fmax(v[0], rb * pbi); v[2] = fmax(v[2], (pu - zu) * q.pc_inv * pb); }
    }
    const int op[4] = {RED_MAX, RED_MAX, RED_MAX, RED_MAX};
    block_reduce<4>(v, op, g_red);
    ResR o;
    const double ps = primal_res_rel > 0 ? primal_res / primal_res_rel : 1.0;
    const double ds = dual_res_rel > 0 ? dual_res / dual_res_rel : 1.0;
    o.primal_res_reg = v[0]; o.primal_res_reg_rel = v[0] / ps;
    o.dual_res_reg = v[1]; o.dual_res_reg_rel = v[1] / ds;
    o.primal_prox_inf = v[2] * delta; o.dual_prox_inf = v[3] * rho;
    return o;
}

// start point (solver.hpp:504-577)
__global__ void k_start_point(IpDev d) {
    const int b = blockIdx.x;
    if (!d.act[b]) return;
    InstPtr q = inst(d, b);
    VarsB it = vb(d.it, d, b), px = vb(d.prox, d, b);
    IpScalars& s = d.sc[b];
    const double n_fin = s.n_fin;
    if (s.has_ineq) {
        double v[2] = {INFINITY, INFINITY};   // min s, min z
        FOR_T(i, d.m) { v[0] = fmin(v[0], fmin(it.s_l[i], it.s_u[i])); v[1] = fmin(v[1], fmin(it.z_l[i], it.z_u[i])); }
        FOR_T(i, d.n) {
            if (q.hxl[i]) { v[0] = fmin(v[0], it.s_bl[i]); v[1] = fmin(v[1], it.z_bl[i]); }
            if (q.hxu[i]) { v[0] = fmin(v[0], it.s_bu[i]); v[1] = fmin(v[1], it.z_bu[i]); }
        }
        const int op[2] = {RED_MIN, RED_MIN};
        block_reduce<2>(v, op, g_red);
        const double ds = fmax(0.0, -v[0]), dz = fmax(0.0, -v[1]);
        FOR_T(i, d.m) {
            if (q.hhl[i]) { it.s_l[i] += ds; it.z_l[i] += dz; }
            if (q.hhu[i]) { it.s_u[i] += ds; it.z_u[i] += dz; }
        }
        FOR_T(i, d.n) {
            if (q.hxl[i]) { it.s_bl[i] += ds; it.z_bl[i] += dz; }
            if (q.hxu[i]) { it.s_bu[i] += ds; it.z_bu[i] += dz; }
        }
        __syncthreads();
        const double mu = fmax(calc_mu(d, q, it, n_fin), 1e-10);
        __syncthreads();
#define FIX(z, sl) { const double c_ = (z) - dz; (z) = (c_ + sqrt(c_ * c_ + 4 * mu)) / 2; (sl) = (z) - c_; }
        FOR_T(i, d.m) { if (q.hhl[i]) FIX(it.z_l[i], it.s_l[i]); if (q.hhu[i]) FIX(it.z_u[i], it.s_u[i]); }
        FOR_T(i, d.n) { if (q.hxl[i]) FIX(it.z_bl[i], it.s_bl[i]); if (q.hxu[i]) FIX(it.z_bu[i], it.s_bu[i]); }
#undef FIX
        __syncthreads();
        const double mu2 = calc_mu(d, q, it, n_fin);
        if (threadIdx.x == 0) s.mu = mu2;
    }
    FOR_T(i, d.n) { px.x[i] = it.x[i]; px.z_bl[i] = it.z_bl[i]; px.z_bu[i] = it.z_bu[i]; }
    FOR_T(i, d.p) px.y[i] = it.y[i];
    FOR_T(i, d.m) { px.z_l[i] = it.z_l[i]; px.z_u[i] = it.z_u[i]; }
}

// loop head (solver.hpp:588-681): termination, regularised residuals, infeasibility, boundary shift, reg-limit finetune
__global__ void k_head(IpDev d) {
    const int b = blockIdx.x;
    if (!d.act[b]) return;
    InstPtr q = inst(d, b);
    VarsB it = vb(d.it, d, b);
    IpScalars& sg = d.sc[b];
    IpScalars s = sg;   // every thread keeps an identical private copy; thread 0 writes it back
    __syncthreads();
    const b200qp_settings& st = d.st;
    if (d.trace && s.iter < d.trace_rows) {
        if (threadIdx.x == 0) {
            double* t = d.trace + ((size_t)b * d.trace_rows + s.iter) * 10;
            t[0] = s.rho; t[1] = s.delta; t[2] = s.mu; t[3] = s.primal_step; t[4] = s.dual_step; t[5] = s.primal_res; t[6] = s.dual_res;
            t[7] = s.primal_obj; t[8] = s.dual_obj; t[9] = s.duality_gap;
        }
    }
    bool stop = false;
    if ((s.primal_res < st.eps_abs || s.primal_res_rel < st.eps_rel) && (s.dual_res < st.eps_abs || s.dual_res_rel < st.eps_rel) &&
        (!st.check_duality_gap || s.duality_gap < st.eps_duality_gap_abs || s.duality_gap_rel < st.eps_duality_gap_rel)) {
        s.status = ST_SOLVED; stop = true;
    }
    if (!stop && (s.primal_res != s.primal_res || s.dual_res != s.dual_res)) { s.status = ST_NUMERICS; stop = true; }   // NaN residuals (k_resid_nr)
    if (!stop) {
        ResR rr = residuals_r(d, b, q, s.rho, s.delta, s.primal_res, s.primal_res_rel, s.dual_res, s.dual_res_rel);
        s.primal_res_reg = rr.primal_res_reg; s.primal_res_reg_rel = rr.primal_res_reg_rel; s.dual_res_reg = rr.dual_res_reg;
        s.dual_res_reg_rel = rr.dual_res_reg_rel; s.primal_prox_inf = rr.primal_prox_inf; s.dual_prox_inf = rr.dual_prox_inf;
        if (s.no_dual_update > min(5, st.reg_finetune_dual_update_threshold) && s.primal_prox_inf > st.infeasibility_threshold &&
            (s.primal_res_reg < st.eps_abs || s.primal_res_reg_rel < st.eps_rel)) { s.status = ST_PRIMAL_INF; stop = true; }
        else if (s.no_primal_update > min(5, st.reg_finetune_primal_update_threshold) && s.dual_prox_inf > st.infeasibility_threshold &&
                 (s.dual_res_reg < st.eps_abs || s.dual_res_reg_rel < st.eps_rel)) { s.status = ST_DUAL_INF; stop = true; }
    }
    if (stop) {
        if (threadIdx.x == 0) { s.active = 0; s.need_factor = 0; sg = s; d.act[b] = 0; d.act2[b] = 0; d.need_factor[b] = 0; }
        return;
    }
    s.iter++;
    // boundary shift (:634-666)
    const double eps = 2.220446049250313e-16;
    double v[3] = {0.0, INFINITY, INFINITY};   // shifted flag, min z_bl, min z_bu
    FOR_T(i, d.m) {
        if (q.hhl[i] && it.z_l[i] < eps) { it.z_l[i] += eps; v[0] = 1.0; }
        if (q.hhu[i] && it.z_u[i] < eps) { it.z_u[i] += eps; v[0] = 1.0; }
    }
    FOR_T(i, d.n) {
        const int fl = q.hxl[i], fu = q.hxu[i];
        const double zl = it.z_bl[i], zu = it.z_bu[i];
        if (fl) v[1] = fmin(v[1], zl);
        if (fu) v[2] = fmin(v[2], zu);
    }
    const int op[3] = {RED_MAX, RED_MIN, RED_MIN};
    block_reduce<3>(v, op, g_red);
    bool shifted = v[0] > 0.0;
    if (v[1] < eps) { FOR_T(i, d.n) if (q.hxl[i]) it.z_bl[i] += eps; shifted = true; }
    if (v[2] < eps) { FOR_T(i, d.n) if (q.hxu[i]) it.z_bu[i] += eps; shifted = true; }
    if (shifted) { __syncthreads(); s.mu = calc_mu(d, q, it, s.n_fin); }
    // reg-limit finetune (:668-681)
    if ((s.no_primal_update > st.reg_finetune_primal_update_threshold && s.rho == s.reg_limit && s.reg_limit != st.reg_finetune_lower_limit) ||
        (s.no_dual_update > st.reg_finetune_dual_update_threshold && s.delta == s.reg_limit && s.reg_limit != st.reg_finetune_lower_limit)) {
        if (s.dual_prox_inf < st.infeasibility_threshold && s.primal_prox_inf < st.infeasibility_threshold) {
            s.reg_limit = st.reg_finetune_lower_limit; s.no_primal_update = 0; s.no_dual_update = 0;
        }
    }
    s.need_factor = 1; s.reg_changed = 0;
    if (threadIdx.x == 0) { sg = s; d.need_factor[b] = 1; }
}

// after the factor loop: regularised residuals again if rho/delta changed (:716-718); predictor rhs (:723-726)
__global__ void k_predictor(IpDev d) {
    const int b = blockIdx.x;
    if (!d.act[b]) return;
    InstPtr q = inst(d, b);
    VarsB it = vb(d.it, d, b), r = vb(d.r, d, b);
    IpScalars& sg = d.sc[b];
    IpScalars s = sg;
    __syncthreads();
    if (s.reg_changed) {
        ResR rr = residuals_r(d, b, q, s.rho, s.delta, s.primal_res, s.primal_res_rel, s.dual_res, s.dual_res_rel);
        if (threadIdx.x == 0) {
            sg.primal_res_reg = rr.primal_res_reg; sg.primal_res_reg_rel = rr.primal_res_reg_rel; sg.dual_res_reg = rr.dual_res_reg;
            sg.dual_res_reg_rel = rr.dual_res_reg_rel; sg.primal_prox_inf = rr.primal_prox_inf; sg.dual_prox_inf = rr.dual_prox_inf;
        }
    }
    if (s.has_ineq) {
        FOR_T(i, d.m) { r.s_l[i] = -it.s_l[i] * it.z_l[i]; r.s_u[i] = -it.s_u[i] * it.z_u[i]; }
        FOR_T(i, d.n) {
            const int fl = q.hxl[i], fu = q.hxu[i];
            const double sl = it.s_bl[i], zl = it.z_bl[i], su = it.s_bu[i], zu = it.z_bu[i];
            if (fl) r.s_bl[i] = -sl * zl;
            if (fu) r.s_bu[i] = -su * zu;
        }
    }
}

// after the predictor solve: step length, sigma, corrector rhs (:739-759); or the full step for problems without inequalities (:844-847)
__global__ void k_corrector(IpDev d) {
    const int b = blockIdx.x;
    if (!d.act[b]) return;
    InstPtr q = inst(d, b);
    VarsB it = vb(d.it, d, b), r = vb(d.r, d, b), sp = vb(d.step, d, b);
    IpScalars& sg = d.sc[b];
    const int has_ineq = sg.has_ineq;
    const double mu = sg.mu, n_fin = sg.n_fin;
    if (!has_ineq) {
        FOR_T(i, d.n) it.x[i] += 1.0 * sp.x[i];
        FOR_T(i, d.p) it.y[i] += 1.0 * sp.y[i];
        if (threadIdx.x == 0) { sg.primal_step = 1.0; sg.dual_step = 1.0; d.act2[b] = 0; }
        return;
    }
    double as, az;
    calc_step(d, q, it, sp, as, az);
    as *= d.st.tau; az *= d.st.tau;
    double v[4] = {0, 0, 0, 0};
    FOR_T(i, d.m) {
        v[0] += (it.s_l[i] + as * sp.s_l[i]) * (it.z_l[i] + az * sp.z_l[i]);
        v[1] += (it.s_u[i] + as * sp.s_u[i]) * (it.z_u[i] + az * sp.z_u[i]);
    }
    FOR_T(i, d.n) {
        const int fl = q.hxl[i], fu = q.hxu[i];
        const double sl = it.s_bl[i], dsl = sp.s_bl[i], zl = it.z_bl[i], dzl = sp.z_bl[i], su = it.s_bu[i], dsu = sp.s_bu[i], zu = it.z_bu[i], dzu = sp.z_bu[i];
        if (fl) v[2] += (sl + as * dsl) * (zl + az * dzl);
        if (fu) v[3] += (su + as * dsu) * (zu + az * dzu);
    }
    const int op[4] = {RED_SUM, RED_SUM, RED_SUM, RED_SUM};
    block_reduce<4>(v, op, g_red);
    double sgm = v[0]; sgm += v[1]; sgm += v[2]; sgm += v[3];
    sgm /= (mu * n_fin);
    sgm = fmax(0.0, fmin(1.0, sgm));
    const double sigma = sgm * sgm * sgm;
    const double sm = sigma * mu;
    FOR_T(i, d.m) { r.s_l[i] += -sp.s_l[i] * sp.z_l[i] + sm; r.s_u[i] += -sp.s_u[i] * sp.z_u[i] + sm; }
    FOR_T(i, d.n) {
        const int fl = q.hxl[i], fu = q.hxu[i];
        const double rl = r.s_bl[i], dsl = sp.s_bl[i], dzl = sp.z_bl[i], ru = r.s_bu[i], dsu = sp.s_bu[i], dzu = sp.z_bu[i];
        if (fl) r.s_bl[i] = rl + (-dsl * dzl + sm);
        if (fu) r.s_bu[i] = ru + (-dsu * dzu + sm);
    }
    if (threadIdx.x == 0) { sg.sigma = sigma; d.act2[b] = 1; }
}

// after the corrector solve: step, iterate update, mu (:771-792)
__global__ void k_update(IpDev d) {
    const int b = blockIdx.x;
    if (!d.act2[b]) return;
    InstPtr q = inst(d, b);
    VarsB it = vb(d.it, d, b), sp = vb(d.step, d, b);
    IpScalars& sg = d.sc[b];
    const double n_fin = sg.n_fin, mu_prev = sg.mu;
    double as, az;
    calc_step(d, q, it, sp, as, az);
    const double ps = as * d.st.tau, dsz = az * d.st.tau;
    FOR_T(i, d.n) {
        const int fl = q.hxl[i], fu = q.hxu[i];
        const double zl = it.z_bl[i], dzl = sp.z_bl[i], sl = it.s_bl[i], dsl = sp.s_bl[i], zu = it.z_bu[i], dzu = sp.z_bu[i], su = it.s_bu[i], dsu = sp.s_bu[i];
        it.x[i] += ps * sp.x[i];
        if (fl) { it.z_bl[i] = zl + dsz * dzl; it.s_bl[i] = sl + ps * dsl; }
        if (fu) { it.z_bu[i] = zu + dsz * dzu; it.s_bu[i] = su + ps * dsu; }
    }
    FOR_T(i, d.p) it.y[i] += dsz * sp.y[i];
    FOR_T(i, d.m) {
        it.z_l[i] += dsz * sp.z_l[i]; it.z_u[i] += dsz * sp.z_u[i];
        it.s_l[i] += ps * sp.s_l[i]; it.s_u[i] += ps * sp.s_u[i];
    }
    __syncthreads();
    const double mu = calc_mu(d, q, it, n_fin);
    if (threadIdx.x == 0) { sg.primal_step = ps; sg.dual_step = dsz; sg.mu = mu; sg.mu_rate = fmax(0.0, (mu_prev - mu) / mu_prev); }
}

// update_residuals_nr (solver.hpp:960-1105): step 1, before the mat-vecs
__global__ void k_resid_pre(IpDev d, const int* mask) {
    const int b = blockIdx.x;
    if (!mask[b]) return;
    VarsB it = vb(d.it, d, b);
    double* wz = PB(d.work_z, d.m);
    FOR_T(i, d.m) wz[i] = it.z_u[i] - it.z_l[i];
}
// step 2, after: rnr.y = -A x, work_x = A^T y, rnr.z_l = G x, work_x2 = G^T (z_u - z_l), rnr.x = -P x
__global__ void k_resid_nr(IpDev d, const int* mask, int first) {
    const int b = blockIdx.x;
    if (!mask[b]) return;
    InstPtr q = inst(d, b);
    VarsB it = vb(d.it, d, b), rnr = vb(d.rnr, d, b);
    IpScalars& sg = d.sc[b];
    double *wx = PB(d.work_x, d.n), *wx2 = PB(d.work_x2, d.n);
    // v: 0 x'Px  1 c'x  2 b'y  3 -h_l'z_l  4 h_u'z_u  5 -x_l'z_bl  6 x_u'z_bu | 7 dual_rel  8 prim_rel  9 primal_res  10 dual_res
    // 11: non-finite flag (fmax drops NaNs, so a NaN iterate would otherwise report residuals of 0)
    double v[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    FOR_T(i, d.n) {
        // every load of the iteration up front (the values behind the flags are only USED under them): a load issued inside a
        // data-dependent branch waits for the flag's own round trip first, and this kernel is nothing but such round trips
        const int fl = q.hxl[i], fu = q.hxu[i];
        const double wa = wx[i], wb = wx2[i], mPx = rnr.x[i], xi = it.x[i], ci = q.c[i], xb = q.xbs[i];
        const double zbl = it.z_bl[i], zbu = it.z_bu[i], sbl = it.s_bl[i], sbu = it.s_bu[i], xl = q.x_l[i], xu = q.x_u[i];
        const double pdi = q.pd_inv[i], pbi = q.pdb_inv[i];      // US_DRES(v, i) = (v * pc_inv) * pd_inv[i], in this order
        double wxi = wa + wb;
        v[7] = fmax(v[7], fabs(mPx * q.pc_inv * pdi));
        v[0] += xi * mPx;            // tmp = -x.(-Px) handled below
        v[1] += ci * xi;
        double rx = mPx - ci;
        v[7] = fmax(v[7], fabs(ci * q.pc_inv * pdi));
        if (fl) { wxi -= xb * zbl; v[5] += xl * zbl; }
        if (fu) { wxi += xb * zbu; v[6] += xu * zbu; }
        v[7] = fmax(v[7], fabs(wxi * q.pc_inv * pdi));
        rx -= wxi;
        rnr.x[i] = rx;
        v[10] = fmax(v[10], fabs(rx * q.pc_inv * pdi));
        if (!isfinite(rx)) v[11] = 1.0;
        // box primal residuals (signed maxima, solver.hpp:1077-1095, 1137-1144)
        if (fl) {
            const double t = xb * xi;
            v[8] = fmax(v[8], t * pbi); v[8] = fmax(v[8], xl * pbi); v[8] = fmax(v[8], sbl * pbi);
            const double rb = t + (-xl - sbl);
            rnr.z_bl[i] = rb;
        }
        if (fu) {
            const double t = -xb * xi;
            v[8] = fmax(v[8], t * pbi); v[8] = fmax(v[8], xu * pbi); v[8] = fmax(v[8], sbu * pbi);
            const double rb = t + (xu - sbu);
            rnr.z_bu[i] = rb;
        }
    }
    FOR_T(i, d.p) {
        const double mAx = rnr.y[i];
        v[8] = fmax(v[8], fabs(US_PRES_EQ(mAx, i)));
        v[2] += q.bv[i] * it.y[i];
        const double ry = mAx + q.bv[i];
        rnr.y[i] = ry;
        v[8] = fmax(v[8], fabs(US_PRES_EQ(q.bv[i], i)));
        v[9] = fmax(v[9], fabs(US_PRES_EQ(ry, i)));
        if (!isfinite(ry)) v[11] = 1.0;
    }
    FOR_T(i, d.m) {
        const double Gx = rnr.z_l[i];
        v[3] += q.h_l[i] * it.z_l[i];
        v[4] += q.h_u[i] * it.z_u[i];
        double rl = 0.0, ru = 0.0;
        if (q.hhl[i]) {
            v[8] = fmax(v[8], US_PRES_INEQ(Gx, i));
            rl = Gx + (-q.h_l[i] - it.s_l[i]);
            v[8] = fmax(v[8], US_PRES_INEQ(q.h_l[i], i)); v[8] = fmax(v[8], US_PRES_INEQ(it.s_l[i], i));
        }
        if (q.hhu[i]) {
            v[8] = fmax(v[8], US_PRES_INEQ(-Gx, i));
            ru = -Gx + (q.h_u[i] - it.s_u[i]);
            v[8] = fmax(v[8], US_PRES_INEQ(q.h_u[i], i)); v[8] = fmax(v[8], US_PRES_INEQ(it.s_u[i], i));
        }
        rnr.z_l[i] = rl; rnr.z_u[i] = ru;
        v[9] = fmax(v[9], fmax(fabs(US_PRES_INEQ(rl, i)), fabs(US_PRES_INEQ(ru, i))));
        if (!isfinite(rl) || !isfinite(ru)) v[11] = 1.0;
    }
    __syncthreads();
    FOR_T(i, d.n) {   // box part of primal_res_nr: signed (solver.hpp:1137-1144)
        if (q.hxl[i]) { v[9] = fmax(v[9], US_PRES_B(rnr.z_bl[i], i)); if (!isfinite(rnr.z_bl[i])) v[11] = 1.0; }
        if (q.hxu[i]) { v[9] = fmax(v[9], US_PRES_B(rnr.z_bu[i], i)); if (!isfinite(rnr.z_bu[i])) v[11] = 1.0; }
    }
    const int op[12] = {RED_SUM, RED_SUM, RED_SUM, RED_SUM, RED_SUM, RED_SUM, RED_SUM, RED_MAX, RED_MAX, RED_MAX, RED_MAX, RED_MAX};
    block_reduce<12>(v, op, g_red);
    if (threadIdx.x == 0) {
        IpScalars s = sg;
        double tmp = -v[0];                       // x'Px
        double pobj = 0.5 * tmp, dobj = -0.5 * tmp;
        double gap_rel = q.pc_inv * fabs(tmp);
        tmp = v[1]; pobj += tmp; gap_rel = fmax(gap_rel, q.pc_inv * fabs(tmp));
        tmp = v[2]; dobj -= tmp; gap_rel = fmax(gap_rel, q.pc_inv * fabs(tmp));
        tmp = -v[3]; dobj -= tmp; gap_rel = fmax(gap_rel, q.pc_inv * fabs(tmp));
        tmp = v[4]; dobj -= tmp; gap_rel = fmax(gap_rel, q.pc_inv * fabs(tmp));
        tmp = -v[5]; dobj -= tmp; gap_rel = fmax(gap_rel, q.pc_inv * fabs(tmp));
        tmp = v[6]; dobj -= tmp; gap_rel = fmax(gap_rel, q.pc_inv * fabs(tmp));
        double gap = fabs(pobj - dobj);
        s.primal_obj = q.pc_inv * pobj; s.dual_obj = q.pc_inv * dobj; s.duality_gap = q.pc_inv * gap;
        s.duality_gap_rel = s.duality_gap / fmax(1.0, gap_rel);
        s.prev_primal_res = s.primal_res; s.prev_dual_res = s.dual_res;
        s.primal_res = v[9]; s.primal_res_rel = v[9] / fmax(1.0, v[8]);
        s.dual_res = v[10]; s.dual_res_rel = v[10] / fmax(1.0, v[7]);
        if (v[11] > 0.0) {   // a non-finite iterate: report it (an Eigen inf-norm of the reference would carry the NaN), k_head stops with PIQP_NUMERICS
            const double qnan = __longlong_as_double(0x7ff8000000000000LL);
            s.primal_res = s.primal_res_rel = s.dual_res = s.dual_res_rel = qnan;
        }
        if (first) { s.prev_primal_res = s.primal_res; s.prev_dual_res = s.dual_res; }
        sg = s;
    }
}

// regularisation update rules (solver.hpp:797-829 with inequalities, :852-876 without)
__global__ void k_reg_update(IpDev d) {
    const int b = blockIdx.x;
    if (!d.act[b]) return;
    VarsB it = vb(d.it, d, b), px = vb(d.prox, d, b);
    InstPtr q = inst(d, b);
    IpScalars& sg = d.sc[b];
    IpScalars s = sg;
    __syncthreads();
    const b200qp_settings& st = d.st;
    const double mu_rate = s.mu_rate;
    bool upd_x, upd_d;
    if (s.has_ineq) {
        upd_x = s.dual_res < 0.95 * s.prev_dual_res || (s.dual_res < st.eps_abs || s.dual_res_rel < st.eps_rel) ||
                (s.rho == st.reg_finetune_lower_limit && s.dual_prox_inf < st.infeasibility_threshold);
        if (upd_x) s.rho = fmax(s.reg_limit, (1.0 - mu_rate) * s.rho);
        else { s.no_primal_update++; if (s.iter < 5 || s.dual_prox_inf < st.infeasibility_threshold) s.rho = fmax(s.reg_limit, (1.0 - 0.666 * mu_rate) * s.rho); }
        upd_d = s.primal_res < 0.95 * s.prev_primal_res || (s.primal_res < st.eps_abs || s.primal_res_rel < st.eps_rel) ||
                (s.delta == st.reg_finetune_lower_limit && s.primal_prox_inf < st.infeasibility_threshold);
        if (upd_d) s.delta = fmax(s.reg_limit, (1.0 - mu_rate) * s.delta);
        else { s.no_dual_update++; if (s.iter < 5 || s.primal_prox_inf < st.infeasibility_threshold) s.delta = fmax(s.reg_limit, (1.0 - 0.666 * mu_rate) * s.delta); }
    } else {
        upd_x = s.dual_res < 0.95 * s.prev_dual_res || (s.dual_res < st.eps_abs || s.dual_res_rel < st.eps_rel);
        if (upd_x) s.rho = fmax(s.reg_limit, 0.1 * s.rho);
        else { s.no_primal_update++; if (s.iter < 5 || s.dual_prox_inf < st.infeasibility_threshold) s.rho = fmax(s.reg_limit, 0.5 * s.rho); }
        upd_d = s.primal_res < 0.95 * s.prev_primal_res || (s.primal_res < st.eps_abs || s.primal_res_rel < st.eps_rel);
        if (upd_d) s.delta = fmax(s.reg_limit, 0.1 * s.delta);
        else { s.no_dual_update++; if (s.iter < 5 || s.primal_prox_inf < st.infeasibility_threshold) s.delta = fmax(s.reg_limit, 0.5 * s.delta); }
    }
    if (upd_x) FOR_T(i, d.n) px.x[i] = it.x[i];
    if (upd_d) {
        FOR_T(i, d.p) px.y[i] = it.y[i];
        if (s.has_ineq) {
            FOR_T(i, d.m) { px.z_l[i] = it.z_l[i]; px.z_u[i] = it.z_u[i]; }
            FOR_T(i, d.n) { const int fl = q.hxl[i], fu = q.hxu[i]; const double zl = it.z_bl[i], zu = it.z_bu[i]; if (fl) px.z_bl[i] = zl; if (fu) px.z_bu[i] = zu; }
        }
    }
    if (threadIdx.x == 0) sg = s;
}

__global__ void k_mark_max_iter(IpDev d) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < d.batch && d.act[b]) { d.sc[b].status = ST_MAX_ITER; d.sc[b].active = 0; d.act[b] = 0; }
}

// unscale_results + restore_dual (solver.hpp:1205-1259); box blocks are already x-indexed
__global__ void k_finish(IpDev d) {
    const int b = blockIdx.x;
    InstPtr q = inst(d, b);
    VarsB it = vb(d.it, d, b);
    FOR_T(i, d.n) {
        it.x[i] = US_PRIMAL(it.x[i], i);
        if (q.hxl[i]) { it.z_bl[i] = US_DUAL_B(it.z_bl[i], i); it.s_bl[i] = US_SLACK_B(it.s_bl[i], i); } else { it.z_bl[i] = 0.0; it.s_bl[i] = kInf; }
        if (q.hxu[i]) { it.z_bu[i] = US_DUAL_B(it.z_bu[i], i); it.s_bu[i] = US_SLACK_B(it.s_bu[i], i); } else { it.z_bu[i] = 0.0; it.s_bu[i] = kInf; }
    }
    FOR_T(i, d.p) it.y[i] = US_DUAL_EQ(it.y[i], i);
    FOR_T(i, d.m) {
        const double zl = US_DUAL_INEQ(it.z_l[i], i), zu = US_DUAL_INEQ(it.z_u[i], i);
        it.z_l[i] = zl; it.z_u[i] = zu;
        it.s_l[i] = zl == 0.0 ? kInf : US_SLACK_INEQ(it.s_l[i], i);
        it.s_u[i] = zu == 0.0 ? kInf : US_SLACK_INEQ(it.s_u[i], i);
    }
}

// =====================================================================================================
// host side
// =====================================================================================================
// The ~90 vectors of the IP loop are carved out of a few zero-filled slabs: every DevBuf::alloc is a cudaMallocAsync plus a stream
// synchronisation (~50 us), which made "allocate the solver" 5-7 ms of a 13 ms end-to-end step of BASELINE config 4.
double* BatchedIPSolver::alloc_d(size_t cnt) {
    const size_t need = (std::max<size_t>(cnt, 1) + 31) & ~(size_t)31;          // 256-byte granules
    if (arena_left_ < need) {
        const size_t chunk = std::max(need, (size_t)24 * (size_t)batch * (size_t)(n + p + m + 8));
        pool_.emplace_back(chunk); pool_.back().zero(stream);
        arena_ptr_ = pool_.back().get(); arena_left_ = chunk;
    }
    double* r = arena_ptr_;
    arena_ptr_ += need; arena_left_ -= need;
    return r;
}
int* BatchedIPSolver::alloc_i(size_t cnt) { return reinterpret_cast<int*>(alloc_d((std::max<size_t>(cnt, 1) + 1) / 2)); }
Vars BatchedIPSolver::alloc_vars() {
    Vars v;
    const size_t B = batch;
    v.x = alloc_d(B * n); v.y = alloc_d(B * p); v.z_l = alloc_d(B * m); v.z_u = alloc_d(B * m); v.z_bl = alloc_d(B * n); v.z_bu = alloc_d(B * n);
    v.s_l = alloc_d(B * m); v.s_u = alloc_d(B * m); v.s_bl = alloc_d(B * n); v.s_bu = alloc_d(B * n);
    return v;
}

BatchedIPSolver::BatchedIPSolver(int batch_, int n_, int p_, int m_, const b200qp_settings& st, cudaStream_t stream_)
    : batch(batch_), n(n_), p(p_), m(m_), stream(stream_) {
    pool_.reserve(128); ipool_.reserve(32);
    IpDev& d = d_;
    const size_t B = batch, N = (size_t)n + p + m;
    d.batch = batch; d.n = n; d.p = p; d.m = m; d.st = st;
    d.c = alloc_d(B * n); d.b = alloc_d(B * p); d.h_l = alloc_d(B * m); d.h_u = alloc_d(B * m); d.x_l = alloc_d(B * n); d.x_u = alloc_d(B * n); d.xbs = alloc_d(B * n);
    d.has_hl = alloc_i(B * m); d.has_hu = alloc_i(B * m); d.has_xl = alloc_i(B * n); d.has_xu = alloc_i(B * n);
    d.pd = alloc_d(B * N); d.pd_inv = alloc_d(B * N); d.pdb = alloc_d(B * n); d.pdb_inv = alloc_d(B * n); d.pc = alloc_d(B); d.pc_inv = alloc_d(B);
    d.it = alloc_vars(); d.r = alloc_vars(); d.rnr = alloc_vars(); d.step = alloc_vars(); d.prox = alloc_vars();
    d.k_s_l = alloc_d(B * m); d.k_s_u = alloc_d(B * m); d.k_s_bl = alloc_d(B * n); d.k_s_bu = alloc_d(B * n);
    d.k_zl_inv = alloc_d(B * m); d.k_zu_inv = alloc_d(B * m); d.k_zbl_inv = alloc_d(B * n); d.k_zbu_inv = alloc_d(B * n);
    d.x_reg = alloc_d(B * n); d.z_reg = alloc_d(B * m); d.z_reg_ir = alloc_d(B * m); d.rhs_x_bar = alloc_d(B * n); d.rhs_z_bar = alloc_d(B * m); d.lhs_z = alloc_d(B * m);
    d.err_x = alloc_d(B * n); d.err_y = alloc_d(B * p); d.err_z = alloc_d(B * m); d.ref_x = alloc_d(B * n); d.ref_y = alloc_d(B * p); d.ref_z = alloc_d(B * m);
    d.work_x = alloc_d(B * n); d.work_x2 = alloc_d(B * n); d.work_z = alloc_d(B * m); d.P_diag = alloc_d(B * n);
    sc_.alloc(B); sc_.zero(stream); d.sc = sc_.get();
    d.delta_reg = alloc_d(B);
    d.need_factor = alloc_i(3 * B); d.act = d.need_factor + 2 * B;      // [need_factor | use_ir | act]: one read-back serves the retry loop and the termination test
    d.act2 = alloc_i(B); d.ok = alloc_i(B); d.ir_mask = alloc_i(B);
    d.trace = nullptr; d.trace_rows = 0;
    if (st.verbose >= 2) { d.trace_rows = st.max_iter + 1; d.trace = alloc_d(B * d.trace_rows * 10); }
    B200_CUDA(cudaMallocHost(&h_flags_, sizeof(int) * std::max<size_t>(3 * B, 3)));
    for (int k = 0; k < 2; k++) { B200_CUDA(cudaMallocHost(&h_flags2_[k], sizeof(int) * std::max<size_t>(3 * B, 3))); B200_CUDA(cudaEventCreateWithFlags(&flag_ev_[k], cudaEventDisableTiming)); }
    for (auto& e : ev_) B200_CUDA(cudaEventCreate(&e));
    // threads per instance CTA of the O(n+m) kernels: few large instances want more memory-level parallelism per CTA
    ipt_ = ((size_t)batch <= 296 && (size_t)n + m >= 1536) ? 512 : IPT;
    if (const char* e = getenv("B200_IPT")) { const int v = atoi(e); if (v == 128 || v == 256 || v == 512) ipt_ = v; }
    if (const char* e = getenv("B200_NO_GRAPH")) use_graphs_ = e[0] != '1';
}
BatchedIPSolver::~BatchedIPSolver() {
    if (h_flags_) cudaFreeHost(h_flags_);
    for (int k = 0; k < 2; k++) { if (h_flags2_[k]) cudaFreeHost(h_flags2_[k]); if (flag_ev_[k]) cudaEventDestroy(flag_ev_[k]); }
    if (graph_exec_) cudaGraphExecDestroy(graph_exec_);
    for (auto& e : ev_) cudaEventDestroy(e);
    for (auto& e : iter_ev_) cudaEventDestroy(e);
}

void BatchedIPSolver::finish_setup(BatchedKKT* backend) {
    be_ = backend;
    B200_LAUNCH(k_counts, batch, ipt_, 0, stream, d_);
    be_->extract_P_diag(d_.P_diag);
}

int BatchedIPSolver::count_flags(const int* dev_flags, int count) {
    if (count < 0) count = batch;
    B200_CUDA(cudaMemcpyAsync(h_flags_, dev_flags, sizeof(int) * count, cudaMemcpyDeviceToHost, stream));
    B200_CUDA(cudaStreamSynchronize(stream));
    int c = 0;
    for (int i = 0; i < batch; i++) c += h_flags_[i] != 0;
    return c;
}

// one round of the factor ladder: vector part, backend factorisation, retry bookkeeping (no host synchronisation)
void BatchedIPSolver::factor_round() {
    B200_ZONE("piqp::KKTSystem::update_scalings_and_factor");
    B200_LAUNCH(k_prepare_factor, batch, ipt_, 0, stream, d_);
    be_->factor(d_.delta_reg, d_.x_reg, d_.z_reg_ir, d_.need_factor, d_.ok);
    B200_LAUNCH(k_after_factor, ceil_div(batch, 128), 128, 0, stream, d_);
}
// read back [need_factor | use_ir | act] (ONE host synchronisation); returns the number of instances whose factorisation is still pending
int BatchedIPSolver::read_factor_flags(int& active) {
    const int pending = count_flags(d_.need_factor, 3 * batch);
    any_ir_ = false; active = 0;
    for (int i = 0; i < batch; i++) { any_ir_ |= h_flags_[batch + i] != 0; active += h_flags_[2 * batch + i] != 0; }
    return pending;
}
// `first_round_done`: the first round was already enqueued (as the tail of the captured iteration graph)
int BatchedIPSolver::factor_with_retry(bool first_round_done) {
    // at most 1 (enable refinement) + max_factor_retires + 1 rounds
    int active = 0;
    for (int round = 0; round < d_.st.max_factor_retires + 3; round++) {
        if (!(round == 0 && first_round_done)) factor_round();
        if (read_factor_flags(active) == 0) break;
    }
    return active;
}

// everything of one IP iteration after its factorisation (solver.hpp:716-880): predictor / corrector solves, step, residuals, regularisation
void BatchedIPSolver::iteration_body(cudaEvent_t after_solves) {
    IpDev& d = d_;
    B200_LAUNCH(k_predictor, batch, ipt_, 0, stream, d);
    kkt_solve(d.r, d.step, d.act);
    B200_LAUNCH(k_corrector, batch, ipt_, 0, stream, d);
    kkt_solve(d.r, d.step, d.act2);
    if (after_solves) B200_CUDA(cudaEventRecord(after_solves, stream));
    B200_LAUNCH(k_update, batch, ipt_, 0, stream, d);
    residuals_nr(d.act);
    B200_LAUNCH(k_resid_nr, batch, ipt_, 0, stream, d, d.act, 0);
    B200_LAUNCH(k_reg_update, batch, ipt_, 0, stream, d);
}

// CUDA graph of [iteration body ; loop head of the next iteration ; first factor round]: an IP iteration is ~35-60 small kernels
// whose launch latency dominates the latency-bound backends (multistage: ~40 % of a step).  Valid while no instance uses iterative
// refinement (its trip count is decided on the host) -- instances whose factorisation fails are caught by the flag read-back that
// follows every replay and go through the stepwise retry ladder.  Re-captured when the settings change.
bool BatchedIPSolver::ensure_graph() {
    if (graph_exec_) return true;
    if (graph_failed_) return false;
    const unsigned long long l0 = g_launches.load();
    cudaGraph_t g = nullptr;
    std::unique_lock<std::shared_mutex> capture_lock(capture_mutex());      // no device-wide synchronisation from other threads meanwhile
    if (cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); graph_failed_ = true; return false; }
    bool ok = true;
    try {
        iteration_body(nullptr);
        B200_LAUNCH(k_head, batch, ipt_, 0, stream, d_);
        factor_round();
    } catch (const std::exception&) { ok = false; }
    if (cudaStreamEndCapture(stream, &g) != cudaSuccess || !g) { cudaGetLastError(); ok = false; }
    capture_lock.unlock();
    if (ok && cudaGraphInstantiate(&graph_exec_, g, 0) != cudaSuccess) { cudaGetLastError(); graph_exec_ = nullptr; ok = false; }
    if (g) cudaGraphDestroy(g);
    graph_launches_ = g_launches.load() - l0;
    if (!ok) graph_failed_ = true;
    return ok;
}
void BatchedIPSolver::drop_graph() {
    if (graph_exec_) { cudaGraphExecDestroy(graph_exec_); graph_exec_ = nullptr; }
    graph_failed_ = false;
}

void BatchedIPSolver::kkt_solve(const Vars& rhs, const Vars& lhs, const int* mask) {
    B200_ZONE("piqp::KKTSystem::solve");
    IpDev& d = d_;
    B200_LAUNCH(k_solve_pre, batch, ipt_, 0, stream, d, rhs, mask);
    be_->solve(d.rhs_x_bar, rhs.y, d.rhs_z_bar, lhs.x, lhs.y, d.lhs_z, mask);
    // iterative refinement: only instances whose last factorisation enabled it (host knows if there are any)
    if (any_ir_) B200_LAUNCH(k_mask_and_flag, ceil_div(batch, 128), 128, 0, stream, mask, d, d.ir_mask);
    if (any_ir_ && count_flags(d.ir_mask) > 0) {
        int* irm = d.ir_mask;
        int* was = ir_was_;
        be_->eval_P_x(1.0, lhs.x, d.err_x, irm);
        be_->eval_A(1.0, 1.0, lhs.x, lhs.y, d.err_y, d.work_x, irm);
        be_->eval_G(1.0, 1.0, lhs.x, d.lhs_z, d.err_z, d.work_x2, irm);
        B200_LAUNCH(k_ir_err, batch, ipt_, 0, stream, d, rhs, lhs.x, lhs.y, d.lhs_z, 0, 0, irm);
        for (int it = 0; it < d.st.iterative_refinement_max_iter; it++) {
            if (count_flags(irm) == 0) break;
            B200_LAUNCH(k_copy_int, ceil_div(batch, 128), 128, 0, stream, irm, was, batch);
            be_->solve(d.err_x, d.err_y, d.err_z, d.ref_x, d.ref_y, d.ref_z, was);
            B200_LAUNCH(k_ir_accum, batch, ipt_, 0, stream, d, lhs.x, lhs.y, d.lhs_z, was);
            be_->eval_P_x(1.0, d.ref_x, d.err_x, was);
            be_->eval_A(1.0, 1.0, d.ref_x, d.ref_y, d.err_y, d.work_x, was);
            be_->eval_G(1.0, 1.0, d.ref_x, d.ref_z, d.err_z, d.work_x2, was);
            B200_LAUNCH(k_ir_err, batch, ipt_, 0, stream, d, rhs, d.ref_x, d.ref_y, d.ref_z, 1, it, was);
            B200_LAUNCH(k_ir_accept, batch, ipt_, 0, stream, d, lhs.x, lhs.y, d.lhs_z, was);
        }
    }
    B200_LAUNCH(k_solve_post, batch, ipt_, 0, stream, d, rhs, lhs, mask);
}

void BatchedIPSolver::residuals_nr(const int* mask) {
    B200_ZONE("piqp::Solver::update_residuals_nr");
    IpDev& d = d_;
    B200_LAUNCH(k_resid_pre, batch, ipt_, 0, stream, d, mask);
    be_->eval_A(-1.0, 1.0, d.it.x, d.it.y, d.rnr.y, d.work_x, mask);
    be_->eval_G(1.0, 1.0, d.it.x, d.work_z, d.rnr.z_l, d.work_x2, mask);
    be_->eval_P_x(-1.0, d.it.x, d.rnr.x, mask);
}

void BatchedIPSolver::solve() {
    B200_ZONE("piqp::Solver::solve");
    IpDev& d = d_;
    const unsigned long long l0 = g_launches.load();
    stats_ = b200qp_stats{};
    float ms;
    B200_CUDA(cudaEventRecord(ev_[4], stream));
    // settings check (solver.hpp:388-392, settings.hpp:84-106)
    const b200qp_settings& st = d.st;
    const bool ok_settings = st.rho_init > 0 && st.delta_init > 0 && st.eps_abs > 0 && st.eps_rel >= 0 && st.eps_duality_gap_abs > 0 &&
        st.eps_duality_gap_rel >= 0 && st.infeasibility_threshold >= 0 && st.reg_lower_limit > 0 && st.reg_finetune_primal_update_threshold >= 0 &&
        st.reg_finetune_dual_update_threshold >= 0 && st.max_iter > 0 && st.max_factor_retires > 0 && st.preconditioner_iter >= 0 && st.tau > 0 &&
        st.tau <= 1 && st.iterative_refinement_eps_abs > 0 && st.iterative_refinement_eps_rel >= 0 && st.iterative_refinement_max_iter >= 0 &&
        st.iterative_refinement_min_improvement_rate >= 1.0 && st.iterative_refinement_static_regularization_eps > 0 &&
        st.iterative_refinement_static_regularization_rel >= 0;
    invalid_settings_ = !ok_settings;
    if (!ok_settings) return;
    if (!ir_was_) ir_was_ = alloc_i(batch);
    any_ir_ = false;

    B200_LAUNCH(k_init, batch, ipt_, 0, stream, d);
    B200_CUDA(cudaEventRecord(ev_[0], stream));
    factor_with_retry(false);
    B200_CUDA(cudaEventRecord(ev_[1], stream));
    B200_LAUNCH(k_initial_rhs, batch, ipt_, 0, stream, d);
    kkt_solve(d.r, d.it, d.act);
    B200_CUDA(cudaEventRecord(ev_[2], stream));
    B200_LAUNCH(k_start_point, batch, ipt_, 0, stream, d);
    residuals_nr(d.act);
    B200_LAUNCH(k_resid_nr, batch, ipt_, 0, stream, d, d.act, 1);
    B200_CUDA(cudaStreamSynchronize(stream));
    B200_CUDA(cudaEventElapsedTime(&ms, ev_[0], ev_[1])); stats_.factor_ms += ms;
    B200_CUDA(cudaEventElapsedTime(&ms, ev_[1], ev_[2])); stats_.solve_ms += ms;

    int L = 0;
    auto iev = [&](int i) { while ((int)iter_ev_.size() <= i) { cudaEvent_t e; B200_CUDA(cudaEventCreate(&e)); iter_ev_.push_back(e); } return iter_ev_[i]; };
    // The reference's per-bucket timers (info.kkt_factor_time / kkt_solve_time, solver.hpp:442-471,484-492) need event records between
    // the phases: with compute_timings or backend profiling on, the iteration is enqueued stepwise; otherwise as one graph replay.
    const bool want_graph = use_graphs_ && be_->graph_capturable() && !be_->profile && !st.compute_timings;
    bucket_timers_ = !want_graph;
    // k_head decides per instance (converged / infeasible / continue) and raises need_factor for the active ones; the masks are
    // read back together with the factorisation flags, so an iteration costs ONE host synchronisation
    B200_LAUNCH(k_head, batch, ipt_, 0, stream, d);
    if (bucket_timers_) B200_CUDA(cudaEventRecord(iev(0), stream));
    int active = factor_with_retry(false);
    while (active > 0) {
        const bool last = L + 1 >= st.max_iter;
        if (want_graph && !any_ir_ && !last && ensure_graph()) {
            if (be_->factor_never_fails()) {
                // The backend never reports a failed factorisation (multistage, like the reference's, multistage_kkt.hpp:218), so nothing
                // the host reads back can change what the NEXT replay does for an active instance: replay L + 1 is enqueued before the
                // flags of replay L are inspected (double-buffered read-back), and the GPU no longer idles for a host round trip per
                // iteration.  A replay that follows the last iteration finds every instance inactive and does nothing.
                int launched = 0, inspected = 0, stop = 0;
                while (!stop) {
                    while (launched - inspected < 2 && L + launched + 1 < st.max_iter) {
                        B200_CUDA(cudaGraphLaunch(graph_exec_, stream));
                        g_launches.fetch_add(graph_launches_, std::memory_order_relaxed);
                        B200_CUDA(cudaMemcpyAsync(h_flags2_[launched & 1], d_.need_factor, sizeof(int) * 3 * batch, cudaMemcpyDeviceToHost, stream));
                        B200_CUDA(cudaEventRecord(flag_ev_[launched & 1], stream));
                        launched++;
                    }
                    if (launched == inspected) break;             // max_iter reached: the stepwise tail below finishes
                    B200_CUDA(cudaEventSynchronize(flag_ev_[inspected & 1]));
                    const int* f = h_flags2_[inspected & 1];
                    int pending = 0; active = 0; bool ir = false;
                    for (int i = 0; i < batch; i++) { pending += f[i] != 0; ir |= f[batch + i] != 0; active += f[2 * batch + i] != 0; }
                    inspected++;
                    if (active == 0 || pending > 0 || ir) stop = 1;
                }
                L += inspected;
                if (launched > inspected) {                       // a speculative replay is in flight: let it drain, its flags are the current ones
                    B200_CUDA(cudaEventSynchronize(flag_ev_[(launched - 1) & 1]));
                    if (active > 0) {                             // stopped for another reason than convergence: the extra replay was a real iteration
                        const int* f = h_flags2_[(launched - 1) & 1];
                        active = 0; any_ir_ = false;
                        for (int i = 0; i < batch; i++) { any_ir_ |= f[batch + i] != 0; active += f[2 * batch + i] != 0; }
                        L += launched - inspected;
                        active = factor_with_retry(true);
                    }
                }
                else if (active > 0 && stop) active = factor_with_retry(true);      // (unreachable for a backend that never fails: pending rounds / refinement)
                if (active == 0) break;
                continue;                                         // max_iter - 1 replays done: the stepwise last iteration follows on the next pass
            }
            B200_CUDA(cudaGraphLaunch(graph_exec_, stream));
            g_launches.fetch_add(graph_launches_, std::memory_order_relaxed);
            L++;
            active = factor_with_retry(true);
            continue;
        }
        if (bucket_timers_) B200_CUDA(cudaEventRecord(iev(3 * L + 1), stream));
        iteration_body(bucket_timers_ ? iev(3 * L + 2) : nullptr);
        L++;
        if (last) break;
        B200_LAUNCH(k_head, batch, ipt_, 0, stream, d);
        if (bucket_timers_) B200_CUDA(cudaEventRecord(iev(3 * L), stream));
        active = factor_with_retry(false);
    }
    stats_.lockstep_iterations = L;
    B200_LAUNCH(k_mark_max_iter, ceil_div(batch, 128), 128, 0, stream, d);
    B200_LAUNCH(k_finish, batch, ipt_, 0, stream, d);
    B200_CUDA(cudaEventRecord(ev_[5], stream));
    B200_CUDA(cudaEventSynchronize(ev_[5]));
    if (bucket_timers_) for (int i = 0; i < L; i++) {
        B200_CUDA(cudaEventElapsedTime(&ms, iter_ev_[3 * i], iter_ev_[3 * i + 1])); stats_.factor_ms += ms;
        B200_CUDA(cudaEventElapsedTime(&ms, iter_ev_[3 * i + 1], iter_ev_[3 * i + 2])); stats_.solve_ms += ms;
    }
    B200_CUDA(cudaEventElapsedTime(&ms, ev_[4], ev_[5])); stats_.total_ms = ms;
    stats_.kernel_launches = g_launches.load() - l0;
}

std::vector<b200qp_info> BatchedIPSolver::infos() {
    std::vector<IpScalars> h(batch);
    B200_CUDA(cudaMemcpyAsync(h.data(), d_.sc, sizeof(IpScalars) * batch, cudaMemcpyDeviceToHost, stream));
    B200_CUDA(cudaStreamSynchronize(stream));
    std::vector<b200qp_info> out(batch);
    stats_.factor_calls = stats_.kkt_solve_calls = stats_.backend_solves = stats_.ip_iterations = 0;
    for (int b = 0; b < batch; b++) {
        const IpScalars& s = h[b];
        b200qp_info& o = out[b];
        std::memset(&o, 0, sizeof o);
        o.status = invalid_settings_ ? ST_INVALID : s.status; o.iter = s.iter; o.rho = s.rho; o.delta = s.delta; o.mu = s.mu; o.sigma = s.sigma;
        o.primal_step = s.primal_step; o.dual_step = s.dual_step;
        o.primal_res = s.primal_res; o.primal_res_rel = s.primal_res_rel; o.dual_res = s.dual_res; o.dual_res_rel = s.dual_res_rel;
        o.primal_res_reg = s.primal_res_reg; o.primal_res_reg_rel = s.primal_res_reg_rel; o.dual_res_reg = s.dual_res_reg; o.dual_res_reg_rel = s.dual_res_reg_rel;
        o.primal_prox_inf = s.primal_prox_inf; o.dual_prox_inf = s.dual_prox_inf; o.prev_primal_res = s.prev_primal_res; o.prev_dual_res = s.prev_dual_res;
        o.primal_obj = s.primal_obj; o.dual_obj = s.dual_obj; o.duality_gap = s.duality_gap; o.duality_gap_rel = s.duality_gap_rel;
        o.factor_retires = s.factor_retires; o.reg_limit = s.reg_limit; o.no_primal_update = s.no_primal_update; o.no_dual_update = s.no_dual_update;
        o.solve_time = stats_.total_ms * 1e-3; o.kkt_factor_time = stats_.factor_ms * 1e-3; o.kkt_solve_time = stats_.solve_ms * 1e-3;
        stats_.factor_calls += s.n_factor; stats_.kkt_solve_calls += s.n_solve; stats_.backend_solves += s.n_backend_solve; stats_.ip_iterations += s.iter;
    }
    return out;
}

}  // namespace b200
