// oracle/oracle_dense.hpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement of the dense KKT backend of PIQP v0.6.2:
//   dense::KKT                 include/piqp/dense/kkt.hpp:25-177
//   dense Ruiz matrix sweeps   include/piqp/dense/preconditioner.hpp:81-139
//   Eigen::LLT<Lower>          (Eigen 3.4.0 is NOT in /root/reference; restated from its published
//                               algorithm: blocked right-looking Cholesky, unblocked below n=32,
//                               block = clamp(16*floor(n/128), 8, 128) -- the same rule the in-tree fork
//                               include/piqp/dense/ldlt_no_pivot.hpp:318-323 uses.)
//   dense::LDLTNoPivot         include/piqp/dense/ldlt_no_pivot.hpp:276-355,432-450
// All matrices are column-major, like Eigen's default.
#pragma once
#include "oracle_core.hpp"

namespace oracle {

// ---------------------------------------------------------------------------------------------
// dense factorisations on a column-major n x n array with leading dimension ld (lower triangle)
// ---------------------------------------------------------------------------------------------

// unblocked Cholesky of the lower triangle; returns -1 on success, else the failing column.
inline int chol_unblocked(double* A, int n, int ld) {
    for (int k = 0; k < n; k++) {
        double x = A[k + (size_t)k * ld];
        for (int j = 0; j < k; j++) x -= A[k + (size_t)j * ld] * A[k + (size_t)j * ld];
        if (x <= 0.0) return k;
        x = std::sqrt(x);
        A[k + (size_t)k * ld] = x;
        const int rs = n - k - 1;
        if (rs > 0) {
            double* col = A + (k + 1) + (size_t)k * ld;
            for (int j = 0; j < k; j++) {
                const double akj = A[k + (size_t)j * ld];
                const double* cj = A + (k + 1) + (size_t)j * ld;
                for (int i = 0; i < rs; i++) col[i] -= cj[i] * akj;
            }
            for (int i = 0; i < rs; i++) col[i] /= x;
        }
    }
    return -1;
}

// blocked right-looking Cholesky (lower).  Returns -1 on success, else failing column.
inline int chol_blocked(double* A, int n, int ld) {
    if (n < 32) return chol_unblocked(A, n, ld);
    int bs = (n / 128) * 16;
    bs = std::max(8, std::min(bs, 128));
    for (int k = 0; k < n; k += bs) {
        const int b = std::min(bs, n - k);
        const int rs = n - k - b;
        double* A11 = A + k + (size_t)k * ld;
        double* A21 = A + (k + b) + (size_t)k * ld;
        double* A22 = A + (k + b) + (size_t)(k + b) * ld;
        int ret = chol_unblocked(A11, b, ld);
        if (ret >= 0) return k + ret;
        if (rs > 0) {
            // A21 <- A21 * L11^{-T}  (column-by-column forward substitution over the b columns)
            for (int j = 0; j < b; j++) {
                double* cj = A21 + (size_t)j * ld;
                for (int l = 0; l < j; l++) {
                    const double ljl = A11[j + (size_t)l * ld];
                    const double* cl = A21 + (size_t)l * ld;
                    for (int i = 0; i < rs; i++) cj[i] -= cl[i] * ljl;
                }
                const double dj = A11[j + (size_t)j * ld];
                for (int i = 0; i < rs; i++) cj[i] /= dj;
            }
            // A22(lower) -= A21 * A21^T
            for (int j = 0; j < rs; j++) {
                double* cj = A22 + j + (size_t)j * ld;
                for (int l = 0; l < b; l++) {
                    const double a = A21[j + (size_t)l * ld];
                    const double* cl = A21 + j + (size_t)l * ld;
                    for (int i = 0; i < rs - j; i++) cj[i] -= cl[i] * a;
                }
            }
        }
    }
    return -1;
}

// x <- L^{-T} L^{-1} x
inline void chol_solve(const double* L, int n, int ld, double* x) {
    for (int j = 0; j < n; j++) {
        x[j] /= L[j + (size_t)j * ld];
        const double xj = x[j];
        const double* c = L + (size_t)j * ld;
        for (int i = j + 1; i < n; i++) x[i] -= c[i] * xj;
    }
    for (int j = n - 1; j >= 0; j--) {
        const double* c = L + (size_t)j * ld;
        double s = x[j];
        for (int i = j + 1; i < n; i++) s -= c[i] * x[i];
        x[j] = s / c[j];
    }
}

// LDLTNoPivot, lower, in place (ldlt_no_pivot.hpp:276-355): unit L below the diagonal, D on the diagonal.
// Uses the unused upper triangle as scratch exactly as the reference does.  Returns -1 on success.
inline int ldlt_unblocked(double* A, int n, int ld, double* temp) {
    for (int k = 0; k < n; k++) {
        const int rs = n - k - 1;
        double* A10 = A + k;  // row k, stride ld, length k
        // temp = A10^T .* D(0..k)
        double x = A[k + (size_t)k * ld];
        if (k > 0) {
            for (int j = 0; j < k; j++) temp[j] = A10[(size_t)j * ld] * A[j + (size_t)j * ld];
            for (int j = 0; j < k; j++) x -= A10[(size_t)j * ld] * temp[j];
            A[k + (size_t)k * ld] = x;
        }
        if (rs > 0) {
            double* col = A + (k + 1) + (size_t)k * ld;
            for (int j = 0; j < k; j++) {
                const double t = temp[j];
                const double* cj = A + (k + 1) + (size_t)j * ld;
                for (int i = 0; i < rs; i++) col[i] -= cj[i] * t;
            }
        }
        if (x == 0.0) return k;   // ldlt_no_pivot.hpp:306
        if (rs > 0) { double* col = A + (k + 1) + (size_t)k * ld; for (int i = 0; i < rs; i++) col[i] /= x; }
    }
    return -1;
}

inline int ldlt_blocked(double* A, int n, int ld, double* temp) {
    if (n < 32) return ldlt_unblocked(A, n, ld, temp);
    int bs = (n / 128) * 16;
    bs = std::max(8, std::min(bs, 128));
    for (int k = 0; k < n; k += bs) {
        const int b = std::min(bs, n - k);
        const int rs = n - k - b;
        double* A11 = A + k + (size_t)k * ld;
        double* A21 = A + (k + b) + (size_t)k * ld;
        double* A22 = A + (k + b) + (size_t)(k + b) * ld;
        int ret = ldlt_unblocked(A11, b, ld, temp);
        if (ret >= 0) return k + ret;
        if (rs > 0) {
            // A21 <- A21 * L11^{-T} (unit lower)
            for (int j = 0; j < b; j++) {
                double* cj = A21 + (size_t)j * ld;
                for (int l = 0; l < j; l++) {
                    const double ljl = A11[j + (size_t)l * ld];
                    const double* cl = A21 + (size_t)l * ld;
                    for (int i = 0; i < rs; i++) cj[i] -= cl[i] * ljl;
                }
            }
            // A21 <- A21 * D11^{-1};  W = A21 * D11;  A22(lower) -= W * A21^T   (ldlt_no_pivot.hpp:345-350)
            std::vector<double> W((size_t)rs * b);
            for (int j = 0; j < b; j++) {
                double* cj = A21 + (size_t)j * ld;
                const double dj = A11[j + (size_t)j * ld];
                const double dinv = 1.0 / dj;
                for (int i = 0; i < rs; i++) { cj[i] *= dinv; W[i + (size_t)j * rs] = cj[i] * dj; }
            }
            for (int j = 0; j < rs; j++) {
                double* cj = A22 + j + (size_t)j * ld;
                for (int l = 0; l < b; l++) {
                    const double a = A21[j + (size_t)l * ld];
                    const double* wl = &W[j + (size_t)l * rs];
                    for (int i = 0; i < rs - j; i++) cj[i] -= wl[i] * a;
                }
            }
        }
    }
    return -1;
}

// x <- L^{-T} D^{-1} L^{-1} x  (ldlt_no_pivot.hpp:432-450)
inline void ldlt_solve(const double* A, int n, int ld, double* x) {
    for (int j = 0; j < n; j++) {
        const double xj = x[j];
        const double* c = A + (size_t)j * ld;
        for (int i = j + 1; i < n; i++) x[i] -= c[i] * xj;
    }
    for (int j = 0; j < n; j++) x[j] /= A[j + (size_t)j * ld];
    for (int j = n - 1; j >= 0; j--) {
        const double* c = A + (size_t)j * ld;
        double s = x[j];
        for (int i = j + 1; i < n; i++) s -= c[i] * x[i];
        x[j] = s;
    }
}

// ---------------------------------------------------------------------------------------------
// dense matrices + backend
// ---------------------------------------------------------------------------------------------
struct DenseMatrices;

// dense/kkt.hpp:25-177
struct DenseKKT : KKTBackend {
    const DenseMatrices& D;
    double m_delta = 0;
    Vec z_reg_inv, kkt, L, AtA, work_z;
    explicit DenseKKT(const DenseMatrices& D_);
    void compute_AtA();
    void update_data(int options) override { if (options & UPDATE_A) compute_AtA(); }
    void assemble(const double* x_reg);                       // update_kkt :140-160
    bool factor(double delta, const double* x_reg, const double* z_reg) override;
    void solve(const double* rx, const double* ry, const double* rz, double* lx, double* ly, double* lz) override;
    void eval_P_x(double alpha, const double* x, double* z) override;
    void eval_A(double an, double at, const double* xn, const double* xt, double* zn, double* zt) override;
    void eval_G(double an, double at, const double* xn, const double* xt, double* zn, double* zt) override;
};

struct DenseMatrices : QPMatrices {
    int n = 0, p = 0, m = 0;
    Vec P, AT, GT;   // P_utri n x n (upper, rest zero); AT n x p; GT n x m ; all column-major
    // optional hook: lets tests plug a foreign backend (the CUDA C-ABI) behind the same caller
    std::unique_ptr<KKTBackend> (*backend_factory)(const DenseMatrices&, void*) = nullptr;
    void* backend_factory_arg = nullptr;

    void resize(int n_, int p_, int m_) { n = n_; p = p_; m = m_; P.assign((size_t)n * n, 0); AT.assign((size_t)n * p, 0); GT.assign((size_t)n * m, 0); }
    // P given column-major full (or upper); keeps the upper triangle only (solver.hpp:182)
    void set_P(const double* Pin) { for (int j = 0; j < n; j++) for (int i = 0; i < n; i++) P[i + (size_t)j * n] = (i <= j) ? Pin[i + (size_t)j * n] : 0.0; }
    void set_AT(const double* ATin) { std::copy(ATin, ATin + (size_t)n * p, AT.begin()); }
    void set_GT(const double* GTin) { std::copy(GTin, GTin + (size_t)n * m, GT.begin()); }

    void set_G_row_zero(int row) override { for (int i = 0; i < n; i++) GT[i + (size_t)row * n] = 0; }

    void kkt_col_norms(const Vec& xbs, Vec& nrm) override {  // dense/preconditioner.hpp:94-110
        for (int k = 0; k < n; k++) {
            double v = 0;
            for (int i = 0; i < k; i++) v = std::max(v, std::fabs(P[i + (size_t)k * n]));
            for (int j = k; j < n; j++) v = std::max(v, std::fabs(P[k + (size_t)j * n]));
            for (int j = 0; j < p; j++) v = std::max(v, std::fabs(AT[k + (size_t)j * n]));
            for (int j = 0; j < m; j++) v = std::max(v, std::fabs(GT[k + (size_t)j * n]));
            nrm[k] = std::max(v, xbs[k]);
        }
        for (int k = 0; k < p; k++) nrm[n + k] = inf_norm(&AT[(size_t)k * n], n);
        for (int k = 0; k < m; k++) nrm[n + p + k] = inf_norm(&GT[(size_t)k * n], n);
    }
    void scale_sym(const double* d) override {  // dense/preconditioner.hpp:118-138
        for (int k = 0; k < n; k++) for (int i = 0; i <= k; i++) P[i + (size_t)k * n] *= d[k];
        for (int k = 0; k < n; k++) for (int j = k; j < n; j++) P[k + (size_t)j * n] *= d[k];
        for (int j = 0; j < p; j++) for (int i = 0; i < n; i++) AT[i + (size_t)j * n] = (d[i] * AT[i + (size_t)j * n]) * d[n + j];
        for (int j = 0; j < m; j++) for (int i = 0; i < n; i++) GT[i + (size_t)j * n] = (d[i] * GT[i + (size_t)j * n]) * d[n + p + j];
    }
    double cost_norm_mean() override {  // dense/preconditioner.hpp:144-151
        double g = 0;
        for (int k = 0; k < n; k++) {
            double v = 0;
            for (int i = 0; i < k; i++) v = std::max(v, std::fabs(P[i + (size_t)k * n]));
            for (int j = k; j < n; j++) v = std::max(v, std::fabs(P[k + (size_t)j * n]));
            g += v;
        }
        return g / double(n);
    }
    void scale_P(double g) override { for (double& e : P) e *= g; }
    void extract_P_diag(Vec& dg) override { for (int i = 0; i < n; i++) dg[i] = P[i + (size_t)i * n]; }
    std::unique_ptr<KKTBackend> make_backend(int) override {
        if (backend_factory) return backend_factory(*this, backend_factory_arg);
        return std::make_unique<DenseKKT>(*this);
    }
};

inline DenseKKT::DenseKKT(const DenseMatrices& D_) : D(D_) {  // dense/kkt.hpp:39-55
    z_reg_inv.assign(D.m, 0); work_z.assign(D.m, 0);
    kkt.assign((size_t)D.n * D.n, 0); L.assign((size_t)D.n * D.n, 0);
    if (D.p > 0) { AtA.assign((size_t)D.n * D.n, 0); compute_AtA(); }
}
inline void DenseKKT::compute_AtA() {  // AT_A(lower) = AT * AT^T
    const int n = D.n, p = D.p;
    if (p == 0) return;
    std::fill(AtA.begin(), AtA.end(), 0.0);
    for (int j = 0; j < n; j++) {
        double* cj = &AtA[j + (size_t)j * n];
        for (int k = 0; k < p; k++) {
            const double a = D.AT[j + (size_t)k * n];
            const double* ck = &D.AT[j + (size_t)k * n];
            for (int i = 0; i < n - j; i++) cj[i] += ck[i] * a;
        }
    }
}
inline void DenseKKT::assemble(const double* x_reg) {
    const int n = D.n, p = D.p, m = D.m;
    for (int j = 0; j < n; j++) for (int i = j; i < n; i++) kkt[i + (size_t)j * n] = D.P[j + (size_t)i * n];
    for (int i = 0; i < n; i++) kkt[i + (size_t)i * n] += x_reg[i];
    if (p > 0) {
        const double s = 1.0 / m_delta;
        for (int j = 0; j < n; j++) for (int i = j; i < n; i++) kkt[i + (size_t)j * n] += s * AtA[i + (size_t)j * n];
    }
    if (m > 0) {
        for (int j = 0; j < n; j++) {
            double* cj = &kkt[j + (size_t)j * n];
            for (int k = 0; k < m; k++) {
                const double w = z_reg_inv[k] * D.GT[j + (size_t)k * n];   // W_delta_inv_G(k, j)
                const double* ck = &D.GT[j + (size_t)k * n];
                for (int i = 0; i < n - j; i++) cj[i] += ck[i] * w;
            }
        }
    }
}
inline bool DenseKKT::factor(double delta, const double* x_reg, const double* z_reg) {  // :73-84
    m_delta = delta;
    for (int i = 0; i < D.m; i++) z_reg_inv[i] = 1.0 / z_reg[i];
    assemble(x_reg);
    L = kkt;
    return chol_blocked(L.data(), D.n, D.n) < 0;
}
inline void DenseKKT::solve(const double* rx, const double* ry, const double* rz, double* lx, double* ly, double* lz) {  // :86-105
    const int n = D.n, p = D.p, m = D.m;
    const double di = 1.0 / m_delta;
    for (int i = 0; i < n; i++) lx[i] = rx[i];
    for (int k = 0; k < m; k++) work_z[k] = z_reg_inv[k] * rz[k];
    for (int k = 0; k < m; k++) { const double w = work_z[k]; const double* c = &D.GT[(size_t)k * n]; for (int i = 0; i < n; i++) lx[i] += c[i] * w; }
    for (int k = 0; k < p; k++) { const double w = di * ry[k]; const double* c = &D.AT[(size_t)k * n]; for (int i = 0; i < n; i++) lx[i] += c[i] * w; }
    chol_solve(L.data(), n, n, lx);
    for (int k = 0; k < p; k++) { ly[k] = di * dot(&D.AT[(size_t)k * n], lx, n); ly[k] -= di * ry[k]; }
    for (int k = 0; k < m; k++) { lz[k] = dot(&D.GT[(size_t)k * n], lx, n); lz[k] -= rz[k]; lz[k] *= z_reg_inv[k]; }
}
inline void DenseKKT::eval_P_x(double alpha, const double* x, double* z) {  // :108-114
    const int n = D.n;
    for (int i = 0; i < n; i++) z[i] = 0;
    for (int j = 0; j < n; j++) {
        const double* c = &D.P[(size_t)j * n];
        const double xj = alpha * x[j];
        double s = 0;
        for (int i = 0; i < j; i++) { z[i] += c[i] * xj; s += c[i] * x[i]; }
        z[j] += c[j] * xj + alpha * s;
    }
}
inline void DenseKKT::eval_A(double an, double at, const double* xn, const double* xt, double* zn, double* zt) {  // :117-123
    const int n = D.n, p = D.p;
    for (int k = 0; k < p; k++) zn[k] = an * dot(&D.AT[(size_t)k * n], xn, n);
    for (int i = 0; i < n; i++) zt[i] = 0;
    for (int k = 0; k < p; k++) { const double w = at * xt[k]; const double* c = &D.AT[(size_t)k * n]; for (int i = 0; i < n; i++) zt[i] += c[i] * w; }
}
inline void DenseKKT::eval_G(double an, double at, const double* xn, const double* xt, double* zn, double* zt) {  // :126-132
    const int n = D.n, m = D.m;
    for (int k = 0; k < m; k++) zn[k] = an * dot(&D.GT[(size_t)k * n], xn, n);
    for (int i = 0; i < n; i++) zt[i] = 0;
    for (int k = 0; k < m; k++) { const double w = at * xt[k]; const double* c = &D.GT[(size_t)k * n]; for (int i = 0; i < n; i++) zt[i] += c[i] * w; }
}

}  // namespace oracle
