// oracle/experimental_multifrontal.hpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE, and not part of the parity oracle either.
//
// A serial restatement of the SUMMATION ORDER of the product's multifrontal LDL^T (piqp_b200/csrc/sparse_frontal.cuh) inside the
// oracle's SparseLDLt, switched on with ORACLE_MULTIFRONTAL=1.  Purpose: to study on the CPU how the iteration path of the
// numerically chaotic Maros-Meszaros problems (QBEACONF, QRECIPE) depends on the order in which the
// Schur-complement contributions are accumulated -- the reference's up-looking algorithm subtracts them one column at a time
// from the running entry, a multifrontal method sums them per subtree first.  Requires a postordered matrix (the product's
// permutation is one).  Same arithmetic per pivot as mf_eliminate_smem: l_i = w_i / d, F(i,c) -= w_i * l_c.
#pragma once
#include <cstdlib>
#include <vector>
#include <algorithm>

namespace oracle {

template <class LDLT, class CscT>
int numeric_multifrontal(LDLT& S, const CscT& A, int variant) {
    const int n = A.rows;
    std::vector<int> cc(n);
    for (int j = 0; j < n; j++) cc[j] = S.Lp[j + 1] - S.Lp[j];
    std::vector<int> sp, sof(n, 0);
    for (int j = 0; j < n; j++) { if (!(j > 0 && S.etree[j - 1] == j && cc[j - 1] == cc[j] + 1)) sp.push_back(j); sof[j] = (int)sp.size() - 1; }
    const int ns = (int)sp.size();
    sp.push_back(n);
    // lower part by columns with values: entry (r, i), r > i, comes from the upper entry (i, r) stored in column r
    std::vector<int> tp(n + 1, 0), tr; std::vector<double> tv, dg(n, 0.0);
    for (int r = 0; r < n; r++) for (int q = A.p[r]; q < A.p[r + 1]; q++) { if (A.i[q] < r) tp[A.i[q] + 1]++; else if (A.i[q] == r) dg[r] = A.x[q]; }
    for (int i = 0; i < n; i++) tp[i + 1] += tp[i];
    tr.assign(tp[n], 0); tv.assign(tp[n], 0.0);
    { std::vector<int> w(tp.begin(), tp.end() - 1); for (int r = 0; r < n; r++) for (int q = A.p[r]; q < A.p[r + 1]; q++) if (A.i[q] < r) { const int t = w[A.i[q]]++; tr[t] = r; tv[t] = A.x[q]; } }
    std::vector<std::vector<int>> kids(ns), U(ns);
    for (int s = 0; s < ns; s++) { const int pj = S.etree[sp[s + 1] - 1]; if (pj >= 0) kids[sof[pj]].push_back(s); }
    std::vector<std::vector<double>> Um(ns);
    std::vector<int> mark(n, -1), pos(n, -1);
    for (int s = 0; s < ns; s++) {
        const int j0 = sp[s], j1 = sp[s + 1] - 1, ws = j1 - j0 + 1;
        std::vector<int>& u = U[s];
        for (int j = j0; j <= j1; j++) for (int t = tp[j]; t < tp[j + 1]; t++) { const int r = tr[t]; if (r > j1 && mark[r] != s) { mark[r] = s; u.push_back(r); } }
        for (int c : kids[s]) for (int r : U[c]) if (r > j1 && mark[r] != s) { mark[r] = s; u.push_back(r); }
        std::sort(u.begin(), u.end());
        if ((int)u.size() != cc[j1]) return -1;
        const int us = (int)u.size(), f = ws + us;
        for (int k = 0; k < ws; k++) pos[j0 + k] = k;
        for (int a = 0; a < us; a++) pos[u[a]] = ws + a;
        std::vector<double> F((size_t)f * f, 0.0);                  // F[i + c * f], lower
        auto add_original = [&]() {
            for (int j = j0; j <= j1; j++) {
                F[(j - j0) + (size_t)(j - j0) * f] += dg[j];
                for (int t = tp[j]; t < tp[j + 1]; t++) F[pos[tr[t]] + (size_t)(j - j0) * f] += tv[t];
            }
        };
        auto add_children = [&]() {
            for (int c : kids[s]) {
                const std::vector<int>& uc = U[c]; const int nc = (int)uc.size();
                const std::vector<double>& M = Um[c];
                for (int b = 0; b < nc; b++) for (int a = b; a < nc; a++) F[pos[uc[a]] + (size_t)pos[uc[b]] * f] += M[a + (size_t)b * nc];
            }
        };
        if (variant == 1) { add_children(); add_original(); } else { add_original(); add_children(); }      // variant 0 = the product's order
        for (int c : kids[s]) { std::vector<double>().swap(Um[c]); }
        for (int k = 0; k < ws; k++) {
            const double d = F[k + (size_t)k * f];
            S.D[j0 + k] = d;
            if (d == 0.0) return j0 + k;
            int q = S.Lp[j0 + k];
            for (int i = k + 1; i < f; i++) { const double l = F[i + (size_t)k * f] / d; S.Li[q] = i < ws ? j0 + i : u[i - ws]; S.Lx[q] = l; q++; }
            for (int c = k + 1; c < f; c++) {
                const double lc = S.Lx[S.Lp[j0 + k] + (c - k - 1)];
                for (int i = c; i < f; i++) F[i + (size_t)c * f] -= F[i + (size_t)k * f] * lc;
            }
            S.Lnz[j0 + k] = f - k - 1;
        }
        Um[s].assign((size_t)us * us, 0.0);
        for (int b = 0; b < us; b++) for (int a = b; a < us; a++) Um[s][a + (size_t)b * us] = F[(ws + a) + (size_t)(ws + b) * f];
    }
    for (int k = 0; k < n; k++) S.Dinv[k] = 1.0 / S.D[k];
    return n;
}

}  // namespace oracle
