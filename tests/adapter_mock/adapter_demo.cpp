// Drives include/piqp_b200_adapter.hpp the way KKTSystem drives a backend (kkt_system.hpp:140-427): through KKTSolverBase pointers.
// Dense, sparse_ldlt (FULL and ALL_ELIMINATED) and multistage adapters on one small block-tridiagonal KKT system; each must solve
//   [[P + diag(x_reg), A^T, G^T], [A, -delta I, 0], [G, 0, -diag(z_reg)]] (x, y, z) = rhs    to 1e-10, clone() must agree bitwise.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include "piqp_b200_adapter.hpp"
using namespace piqp;
static const int n = 12, p = 4, m = 6;
static double Pd[n][n], Ad[p][n], Gd[m][n];
static void build() {
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) Pd[i][j] = i == j ? 4.0 + 0.1 * i : (std::abs(i - j) == 1 ? -1.0 : 0.0);
    for (int i = 0; i < p; i++) for (int j = 0; j < n; j++) Ad[i][j] = (j == 3 * i || j == 3 * i + 1) ? 1.0 + 0.2 * j : 0.0;
    for (int i = 0; i < m; i++) for (int j = 0; j < n; j++) Gd[i][j] = (j == 2 * i || j == 2 * i + 1) ? 0.5 - 0.1 * i : 0.0;
}
template<class M> static void to_csc(SparseMat<double, int>& S, int rows, int cols, M at, bool upper) {
    S.r = rows; S.c = cols; S.outer.assign(1, 0);
    for (int j = 0; j < cols; j++) { for (int i = 0; i < rows; i++) if ((!upper || i <= j) && at(i, j) != 0.0) { S.inner.push_back(i); S.val.push_back(at(i, j)); } S.outer.push_back((int)S.inner.size()); }
}
template<class Solver, class DataT> static double run(Solver* s, const DataT& d, const char* name) {
    Vec<double> xr(n), zr(m), rx(n), ry(p), rz(m), lx(n), ly(p), lz(m);
    for (int i = 0; i < n; i++) { xr(i) = 0.3 + 0.01 * i; rx(i) = std::sin(1.0 + i); }
    for (int i = 0; i < m; i++) { zr(i) = 0.7 + 0.05 * i; rz(i) = std::cos(2.0 + i); }
    for (int i = 0; i < p; i++) ry(i) = 0.1 * (i + 1);
    const double delta = 0.25;
    if (!s->update_scalings_and_factor(d, delta, xr, zr)) { std::printf("%s: factor failed\n", name); return 1e300; }
    s->solve(d, rx, ry, rz, lx, ly, lz);
    double err = 0;
    for (int i = 0; i < n; i++) { double r = xr(i) * lx(i) - rx(i); for (int j = 0; j < n; j++) r += Pd[i][j] * lx(j); for (int k = 0; k < p; k++) r += Ad[k][i] * ly(k); for (int k = 0; k < m; k++) r += Gd[k][i] * lz(k); err = std::fmax(err, std::fabs(r)); }
    for (int k = 0; k < p; k++) { double r = -delta * ly(k) - ry(k); for (int j = 0; j < n; j++) r += Ad[k][j] * lx(j); err = std::fmax(err, std::fabs(r)); }
    for (int k = 0; k < m; k++) { double r = -zr(k) * lz(k) - rz(k); for (int j = 0; j < n; j++) r += Gd[k][j] * lx(j); err = std::fmax(err, std::fabs(r)); }
    auto c = s->clone();
    Vec<double> cx(n), cy(p), cz(m);
    c->solve(d, rx, ry, rz, cx, cy, cz);
    for (int i = 0; i < n; i++) if (cx(i) != lx(i)) err = 1e300;
    Vec<double> z(n), zn(p), zt(n);
    s->eval_P_x(d, 2.0, lx, z);
    for (int i = 0; i < n; i++) { double r = -z(i); for (int j = 0; j < n; j++) r += 2.0 * Pd[i][j] * lx(j); err = std::fmax(err, std::fabs(r)); }
    s->eval_A_xn_and_AT_xt(d, 1.0, -1.0, lx, ly, zn, zt);
    for (int k = 0; k < p; k++) { double r = -zn(k); for (int j = 0; j < n; j++) r += Ad[k][j] * lx(j); err = std::fmax(err, std::fabs(r)); }
    std::printf("%-28s residual %.3e\n", name, err);
    return err;
}
int main() {
    build();
    dense::Data<double> dd; dd.n = n; dd.p = p; dd.m = m;
    dd.P_utri = Mat<double>(n, n); dd.AT = Mat<double>(n, p); dd.GT = Mat<double>(n, m);
    for (int i = 0; i < n; i++) for (int j = i; j < n; j++) dd.P_utri(i, j) = Pd[i][j];
    for (int k = 0; k < p; k++) for (int j = 0; j < n; j++) dd.AT(j, k) = Ad[k][j];
    for (int k = 0; k < m; k++) for (int j = 0; j < n; j++) dd.GT(j, k) = Gd[k][j];
    sparse::Data<double, int> sd; sd.n = n; sd.p = p; sd.m = m;
    to_csc(sd.P_utri, n, n, [](int i, int j) { return Pd[i][j]; }, true);
    to_csc(sd.AT, n, p, [](int i, int k) { return Ad[k][i]; }, false);
    to_csc(sd.GT, n, m, [](int i, int k) { return Gd[k][i]; }, false);
    double worst = 0;
    { std::unique_ptr<KKTSolverBase<double, int, PIQP_DENSE>> s = std::make_unique<b200::DenseKKT<double>>(dd); worst = std::fmax(worst, run(s.get(), dd, "b200::DenseKKT")); }
    { std::unique_ptr<KKTSolverBase<double, int, PIQP_SPARSE>> s = std::make_unique<b200::SparseKKT<double, int, 0>>(sd); worst = std::fmax(worst, run(s.get(), sd, "b200::SparseKKT<KKT_FULL>")); }
    { std::unique_ptr<KKTSolverBase<double, int, PIQP_SPARSE>> s = std::make_unique<b200::SparseKKT<double, int, 3>>(sd); worst = std::fmax(worst, run(s.get(), sd, "b200::SparseKKT<KKT_ALL_ELIM>")); }
    { std::unique_ptr<KKTSolverBase<double, int, PIQP_SPARSE>> s = std::make_unique<b200::MultistageKKT<double, int>>(sd); worst = std::fmax(worst, run(s.get(), sd, "b200::MultistageKKT")); }
    std::printf("%s\n", worst < 1e-10 ? "ADAPTER_OK" : "ADAPTER_FAIL");
    return worst < 1e-10 ? 0 : 1;
}
