// piqp_b200/csrc/sparse_ldlt_backend.cu -- see sparse_ldlt_backend.hpp
#include "sparse_ldlt_backend.hpp"
#include "sparse_frontal.cuh"
#include "sparse_wide.cuh"
#include <cstdlib>
#include <chrono>
#include <string>
#include <algorithm>
#include <cstdio>
#include <set>
#include <stdexcept>

namespace b200 {

// =====================================================================================================
// host: ordering
// =====================================================================================================
// Minimum-degree ordering on the graph of a symmetric matrix given by its upper triangle (CSC).  Quotient graph
// (variables + elements, element absorption) with exact external degrees; ties broken by index.  perm[k] = k-th pivot.
static std::vector<int> minimum_degree_ordering_exact(int n, const std::vector<int>& cp, const std::vector<int>& ri) {
    std::vector<std::vector<int>> avar(n), aelem(n), members(n);
    for (int j = 0; j < n; j++) for (int q = cp[j]; q < cp[j + 1]; q++) { const int i = ri[q]; if (i != j) { avar[i].push_back(j); avar[j].push_back(i); } }
    for (auto& v : avar) { std::sort(v.begin(), v.end()); v.erase(std::unique(v.begin(), v.end()), v.end()); }
    std::vector<char> gone(n, 0), elem_alive(n, 0);
    std::vector<int> deg(n), mark(n, -1), mark2(n, -1);
    std::set<std::pair<int, int>> pq;
    for (int i = 0; i < n; i++) { deg[i] = (int)avar[i].size(); pq.insert({deg[i], i}); }
    std::vector<int> perm; perm.reserve(n);
    std::vector<int> Lp;
    int stamp = 0;
    while (!pq.empty()) {
        const int pv = pq.begin()->second;
        pq.erase(pq.begin());
        perm.push_back(pv);
        // pattern of the new element
        Lp.clear();
        mark[pv] = pv;
        for (int v : avar[pv]) if (!gone[v] && mark[v] != pv) { mark[v] = pv; Lp.push_back(v); }
        for (int e : aelem[pv]) {
            if (!elem_alive[e]) continue;
            for (int v : members[e]) if (!gone[v] && mark[v] != pv) { mark[v] = pv; Lp.push_back(v); }
            elem_alive[e] = 0; members[e].clear(); members[e].shrink_to_fit();
        }
        gone[pv] = 1;
        avar[pv].clear(); aelem[pv].clear();
        members[pv] = Lp; elem_alive[pv] = 1;
        for (int i : Lp) {
            // prune: variables now reachable through the new element, dead elements
            auto& av = avar[i];
            size_t w = 0;
            for (int v : av) if (!gone[v] && mark[v] != pv) av[w++] = v;
            av.resize(w);
            auto& ae = aelem[i];
            w = 0;
            for (int e : ae) if (elem_alive[e] && e != pv) ae[w++] = e;
            ae.resize(w);
            ae.push_back(pv);
            // exact external degree
            ++stamp;
            int d = 0;
            mark2[i] = stamp;
            for (int v : av) if (mark2[v] != stamp) { mark2[v] = stamp; d++; }
            for (int e : ae) for (int v : members[e]) if (!gone[v] && mark2[v] != stamp) { mark2[v] = stamp; d++; }
            pq.erase({deg[i], i});
            deg[i] = d;
            pq.insert({d, i});
        }
    }
    return perm;
}


// Approximate minimum degree on the same quotient graph (what Eigen::AMDOrdering -- third party, ordering.hpp:72-74 --
// computes in spirit: Amestoy/Davis/Duff's degree bound  d_i <= |A_i| + |L_p \ i| + sum_{e in E_i} |L_e \ L_p|  evaluated with
// one pass over the element lists, element absorption, aggressive absorption of elements that became subsets of the new
// one, degree buckets).  No supervariables; instead the elimination stops as soon as the bound says the remaining graph is
// a clique (min degree = remaining - 1): the rest is appended in index order and becomes the dense root front.  Cost
// O(sum_pivots sum_{i in L_p} (|A_i| + |E_i|)): BASELINE config 3 (n_kkt = 20 000, 1 %) takes ~1 s instead of the 94 s of
// the exact-degree version above (kept under B200_ORDERING=exact for cross-checks).
std::vector<int> minimum_degree_ordering(int n, const std::vector<int>& cp, const std::vector<int>& ri) {
    B200_ZONE("piqp::AMDOrdering::init");
    if (const char* e = getenv("B200_ORDERING")) if (std::string(e) == "exact") return minimum_degree_ordering_exact(n, cp, ri);
    std::vector<std::vector<int>> avar(n), aelem(n), members(n);
    // Dense rows / columns (AMD's rule: degree > max(16, 10 sqrt(n))) are taken out of the graph up front and ordered last: a KKT node
    // that touches tens of thousands of variables (BOYD1: 18 equality rows with 31 000 entries each) would otherwise be rescanned at
    // every one of the n pivots that touch it (measured: 19.7 s -> 0.3 s for the ordering of BOYD1, n_kkt = 93 279).
    std::vector<int> deg0(n, 0);
    for (int j = 0; j < n; j++) for (int q = cp[j]; q < cp[j + 1]; q++) { const int i = ri[q]; if (i != j) { deg0[i]++; deg0[j]++; } }
    const int dense_thr = std::max(16, (int)(10.0 * std::sqrt((double)n)));
    std::vector<char> gone(n, 0), elem_alive(n, 0);
    std::vector<int> dense_nodes;
    for (int i = 0; i < n; i++) if (deg0[i] > dense_thr) { gone[i] = 1; dense_nodes.push_back(i); }
    for (int j = 0; j < n; j++) for (int q = cp[j]; q < cp[j + 1]; q++) { const int i = ri[q]; if (i != j && !gone[i] && !gone[j]) { avar[i].push_back(j); avar[j].push_back(i); } }
    for (auto& v : avar) { std::sort(v.begin(), v.end()); v.erase(std::unique(v.begin(), v.end()), v.end()); }
    std::vector<int> deg(n), mark(n, -1), head(n + 1, -1), nxt(n, -1), prv(n, -1);
    std::vector<long long> w(n, 0);
    std::vector<int> msize(n, 0);
    long long wflg = 1;
    const int n_all = n;
    n -= (int)dense_nodes.size();          // the elimination below runs on the n sparse nodes (bucket indices / clique tests use this count)
    auto bucket_insert = [&](int i) { const int d = deg[i]; prv[i] = -1; nxt[i] = head[d]; if (head[d] >= 0) prv[head[d]] = i; head[d] = i; };
    auto bucket_remove = [&](int i) { const int d = deg[i]; if (prv[i] >= 0) nxt[prv[i]] = nxt[i]; else head[d] = nxt[i]; if (nxt[i] >= 0) prv[nxt[i]] = prv[i]; };
    for (int i = n_all - 1; i >= 0; i--) if (!gone[i]) { deg[i] = (int)avar[i].size(); bucket_insert(i); }      // reverse: the smallest index sits at the head
    std::vector<int> perm; perm.reserve(n_all);
    std::vector<int> Lp;
    int mindeg = 0;
    const int dense_pct = getenv("B200_AMD_DENSE_PCT") ? atoi(getenv("B200_AMD_DENSE_PCT")) : 40, dense_abs = 128;
    for (int k = 0; k < n; k++) {
        const int nleft = n - k;
        while (mindeg < n && head[mindeg] < 0) mindeg++;
        // dense tail: the remaining graph is (bounded by) a clique, or every remaining variable already touches >= 40 % of the
        // others (and >= 128): eliminating them one by one would create a handful of huge, nearly identical fronts; one dense
        // root front costs a few percent more flops and runs at tensor-pipe speed
        if (nleft > 1 && (mindeg >= nleft - 1 || (mindeg >= dense_abs && (long long)mindeg * 100 >= (long long)dense_pct * nleft))) {
            for (int i = 0; i < n_all; i++) if (!gone[i]) perm.push_back(i);
            break;
        }
        const int pv = head[mindeg];
        bucket_remove(pv);
        perm.push_back(pv);
        Lp.clear();
        mark[pv] = pv;
        for (int v : avar[pv]) if (!gone[v] && mark[v] != pv) { mark[v] = pv; Lp.push_back(v); }
        for (int e : aelem[pv]) {
            if (!elem_alive[e]) continue;
            for (int v : members[e]) if (!gone[v] && mark[v] != pv) { mark[v] = pv; Lp.push_back(v); }
            elem_alive[e] = 0; std::vector<int>().swap(members[e]);
        }
        gone[pv] = 1;
        std::vector<int>().swap(avar[pv]); std::vector<int>().swap(aelem[pv]);
        const int lp = (int)Lp.size();
        // |L_e \ L_p| for every element adjacent to L_p: w[e] - wflg
        for (int i : Lp) for (int e : aelem[i]) { if (!elem_alive[e]) continue; if (w[e] < wflg) w[e] = (long long)msize[e] + wflg; w[e]--; }      // msize: |L_e| without touching the vector's header
        for (int i : Lp) {
            auto& av = avar[i];
            size_t o = 0;
            for (int v : av) if (!gone[v] && mark[v] != pv) av[o++] = v;
            av.resize(o);
            auto& ae = aelem[i];
            o = 0;
            long long ext = 0;
            for (int e : ae) {
                if (!elem_alive[e]) continue;
                const long long d = w[e] - wflg;
                if (d <= 0) { elem_alive[e] = 0; std::vector<int>().swap(members[e]); continue; }     // aggressive absorption: L_e is a subset of L_p
                ext += d; ae[o++] = e;
            }
            ae.resize(o);
            ae.push_back(pv);
            long long d = (long long)av.size() + (lp - 1) + ext;
            d = std::min<long long>(d, (long long)deg[i] + (lp - 1));
            d = std::min<long long>(d, nleft - 2);               // nleft - 1 variables remain after this pivot, i is one of them
            if (d < 0) d = 0;
            bucket_remove(i);
            deg[i] = (int)d;
            bucket_insert(i);
            if (deg[i] < mindeg) mindeg = deg[i];
        }
        members[pv] = Lp; msize[pv] = lp; elem_alive[pv] = lp > 0;
        wflg += (long long)n_all + 1;
    }
    std::stable_sort(dense_nodes.begin(), dense_nodes.end(), [&](int a, int b) { return deg0[a] < deg0[b]; });
    perm.insert(perm.end(), dense_nodes.begin(), dense_nodes.end());
    return perm;
}

// =====================================================================================================
// host: symbolic analysis
// =====================================================================================================
// Elimination tree and column counts of L for the permuted upper pattern PK (column k: rows i <= k, sorted), in O(nnz(PK) alpha): Liu's
// elimination tree with path compression, then the skeleton / least-common-ancestor column counts of Gilbert, Ng and Peyton (SIAM J. Matrix
// Anal. Appl. 15, 1994; the formulation of Davis, "Direct Methods for Sparse Linear Systems", 4.5).  The reference computes the same two
// arrays by walking every row subtree (sparse/ldlt.hpp:42-99), which is O(nnz(L)): 51 M steps per pass for BASELINE config 3, and the
// symbolic phase needs them twice (before the postorder and after the amalgamation).  Lnz[j] = |struct(L_j)| without the diagonal.
// B200_SYMBOLIC_WALK=1 keeps the reference's walk (tests compare the two).
static void etree_and_column_counts(int nk, const std::vector<int>& PKp, const std::vector<int>& PKi_rows, std::vector<int>& etree, std::vector<int>& Lnz) {
    etree.assign(nk, -1);
    std::vector<int> anc(nk, -1);
    for (int k = 0; k < nk; k++)
        for (int q = PKp[k]; q < PKp[k + 1]; q++) {
            int i = PKi_rows[q];
            while (i != -1 && i < k) { const int nx = anc[i]; anc[i] = k; if (nx == -1) etree[i] = k; i = nx; }
        }
    // a postorder (children by increasing index)
    std::vector<int> head(nk, -1), next(nk, -1), post; post.reserve(nk);
    for (int j = nk - 1; j >= 0; j--) if (etree[j] >= 0) { next[j] = head[etree[j]]; head[etree[j]] = j; }
    {
        std::vector<int> stack;
        for (int r = 0; r < nk; r++) {
            if (etree[r] >= 0) continue;
            stack.push_back(r);
            while (!stack.empty()) {
                const int v = stack.back();
                if (head[v] >= 0) { const int c = head[v]; head[v] = next[c]; stack.push_back(c); }
                else { post.push_back(v); stack.pop_back(); }
            }
        }
    }
    // rows i > j of column j of the LOWER pattern (= the transpose of the strictly upper part of PK)
    std::vector<int> tp(nk + 1, 0), tr;
    for (int k = 0; k < nk; k++) for (int q = PKp[k]; q < PKp[k + 1]; q++) if (PKi_rows[q] < k) tp[PKi_rows[q] + 1]++;
    for (int i = 0; i < nk; i++) tp[i + 1] += tp[i];
    tr.assign(tp[nk], 0);
    { std::vector<int> w(tp.begin(), tp.end() - 1); for (int k = 0; k < nk; k++) for (int q = PKp[k]; q < PKp[k + 1]; q++) if (PKi_rows[q] < k) tr[w[PKi_rows[q]]++] = k; }
    std::vector<int> first(nk, -1), maxfirst(nk, -1), prevleaf(nk, -1), delta(nk, 0);
    for (int k = 0; k < nk; k++) {
        int j = post[k];
        delta[j] = (first[j] == -1) ? 1 : 0;                     // 1 for a leaf of the elimination tree
        for (; j != -1 && first[j] == -1; j = etree[j]) first[j] = k;
    }
    for (int i = 0; i < nk; i++) anc[i] = i;
    for (int k = 0; k < nk; k++) {
        const int j = post[k];
        if (etree[j] != -1) delta[etree[j]]--;
        for (int t = tp[j]; t < tp[j + 1]; t++) {
            const int i = tr[t];                                 // A(i, j) != 0, i > j
            if (first[j] <= maxfirst[i]) continue;               // j is not a leaf of the i-th row subtree
            maxfirst[i] = first[j];
            const int jprev = prevleaf[i];
            prevleaf[i] = j;
            delta[j]++;                                          // A(i, j) is in the skeleton
            if (jprev != -1) {                                   // subsequent leaf: remove the overlap at q = lca(jprev, j)
                int q = jprev;
                while (q != anc[q]) q = anc[q];
                for (int s2 = jprev; s2 != q;) { const int sp = anc[s2]; anc[s2] = q; s2 = sp; }
                delta[q]--;
            }
        }
        if (etree[j] != -1) anc[j] = etree[j];
    }
    for (int j = 0; j < nk; j++) if (etree[j] != -1) delta[etree[j]] += delta[j];
    Lnz.resize(nk);
    for (int j = 0; j < nk; j++) Lnz[j] = delta[j] - 1;
}

// Fundamental supernodes of a postordered symbolic factorisation (chains j -> j+1 of the elimination tree whose column counts
// drop by one) and their update-row sets U_s = struct(L_j1) = rows > j1 of [the matrix columns of s  u  the U's of the child
// supernodes]: the columns of s have the patterns {j+1..j1} u U_s.  One merge per supernode, no walk over nnz(L).
static bool supernode_patterns(int nk, const std::vector<int>& etree, const std::vector<int>& Lnz, const std::vector<int>& PKp, const std::vector<int>& PKi_rows,
                               std::vector<int>& sp, std::vector<int>& sof, std::vector<std::vector<int>>& U, std::string& error) {
    sp.clear(); sof.assign(nk, 0);
    for (int j = 0; j < nk; j++) { if (!(j > 0 && etree[j - 1] == j && Lnz[j - 1] == Lnz[j] + 1)) sp.push_back(j); sof[j] = (int)sp.size() - 1; }
    const int ns = (int)sp.size();
    sp.push_back(nk);
    std::vector<int> tp(nk + 1, 0), tr;                  // lower pattern by columns: rows r > i with PK(i, r) != 0
    for (int r = 0; r < nk; r++) for (int q = PKp[r]; q < PKp[r + 1]; q++) if (PKi_rows[q] < r) tp[PKi_rows[q] + 1]++;
    for (int i = 0; i < nk; i++) tp[i + 1] += tp[i];
    tr.assign(tp[nk], 0);
    { std::vector<int> w(tp.begin(), tp.end() - 1); for (int r = 0; r < nk; r++) for (int q = PKp[r]; q < PKp[r + 1]; q++) if (PKi_rows[q] < r) tr[w[PKi_rows[q]]++] = r; }
    std::vector<int> chead(ns, -1), cnext(ns, -1), mark(nk, -1);
    for (int s2 = ns - 1; s2 >= 0; s2--) { const int pj = etree[sp[s2 + 1] - 1]; if (pj >= 0) { cnext[s2] = chead[sof[pj]]; chead[sof[pj]] = s2; } }
    U.assign(ns, std::vector<int>());
    for (int s2 = 0; s2 < ns; s2++) {
        const int j0 = sp[s2], j1 = sp[s2 + 1] - 1;
        std::vector<int>& u = U[s2];
        for (int j = j0; j <= j1; j++) for (int t = tp[j]; t < tp[j + 1]; t++) { const int r = tr[t]; if (r > j1 && mark[r] != s2) { mark[r] = s2; u.push_back(r); } }
        for (int c = chead[s2]; c >= 0; c = cnext[c]) for (int r : U[c]) if (r > j1 && mark[r] != s2) { mark[r] = s2; u.push_back(r); }
        std::sort(u.begin(), u.end());
        if ((int)u.size() != Lnz[j1]) { error = "sparse_ldlt: internal error (supernodal pattern does not match the column counts)"; return false; }
    }
    return true;
}

// Contribution lists of upper(MT * diag(w) * MT^T) for MT given column-wise (n x r CSC): entry (i, j), i <= j, is
// sum_k MT(j,k) * MT(i,k) * w_k over the columns k that hold both rows, k ascending -- the order in which the reference's
// Gustavson loops accumulate it (kkt_all_eliminated.hpp:178-220).  pa / pb are value indices of MT(i,k) / MT(j,k).
void LdltSymbolic::build_gram(const Pattern& MT, Gram& g) {
    const int nr = MT.rows, nc = MT.cols;
    std::vector<int> rp(nr + 1, 0), rk(MT.nnz), rpos(MT.nnz);          // row view of MT (columns ascending within a row)
    for (int q = 0; q < MT.nnz; q++) rp[MT.i[q] + 1]++;
    for (int r = 0; r < nr; r++) rp[r + 1] += rp[r];
    { std::vector<int> w(rp.begin(), rp.end() - 1); for (int k = 0; k < nc; k++) for (int q = MT.p[k]; q < MT.p[k + 1]; q++) { const int t = w[MT.i[q]]++; rk[t] = k; rpos[t] = q; } }
    g.colp.assign(nr + 1, 0); g.rows.clear(); g.ptr.assign(1, 0); g.pa.clear(); g.pb.clear();
    struct Tr { int i, pa, pb; };
    std::vector<Tr> tr;
    for (int j = 0; j < nr; j++) {
        tr.clear();
        for (int a = rp[j]; a < rp[j + 1]; a++) { const int k = rk[a];
            for (int t = MT.p[k]; t < MT.p[k + 1]; t++) if (MT.i[t] <= j) tr.push_back({MT.i[t], t, rpos[a]}); }
        std::stable_sort(tr.begin(), tr.end(), [](const Tr& x, const Tr& y) { return x.i < y.i; });
        for (size_t t = 0; t < tr.size(); t++) {
            if (t == 0 || tr[t].i != tr[t - 1].i) { g.rows.push_back(tr[t].i); g.ptr.push_back(g.ptr.back()); }
            g.pa.push_back(tr[t].pa); g.pb.push_back(tr[t].pb); g.ptr.back()++;
        }
        g.colp[j + 1] = (int)g.rows.size();
    }
}

bool LdltSymbolic::analyse(const Pattern& P, const Pattern& AT, const Pattern& GT, const int* user_perm, int mode_) {
    B200_ZONE("piqp::LDLt::factorize_symbolic_upper_triangular");
    mode = mode_;
    const bool dbg_t = getenv("B200_DEBUG_SYMBOLIC") != nullptr;
    auto t_last = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) { if (!dbg_t) return; auto t = std::chrono::steady_clock::now(); fprintf(stderr, "  [symbolic %.3f s] %s\n", std::chrono::duration<double>(t - t_last).count(), what); t_last = t; };
    const bool elim_eq = mode & 1, elim_ineq = mode & 2;
    n = P.rows; p = AT.cols; m = GT.cols;
    pk = elim_eq ? 0 : p; mk = elim_ineq ? 0 : m;
    nk = n + pk + mk;
    // ---- KKT pattern, upper CSC.  FULL: [[P + rho I, A^T, G^T], [., -delta I, .], [., ., -Z]] (kkt_full.hpp:39-170); the
    // condensed modes fold delta^-1 A^T A and / or G^T Z^-1 G into the top-left block on the structural union pattern and
    // drop the corresponding block columns (kkt_eq_eliminated.hpp:47-148, kkt_ineq_eliminated.hpp:49-150, kkt_all_eliminated.hpp:58-106)
    Kp.assign(nk + 1, 0); Ki.clear();
    P_to_K.assign(P.nnz, -1); AT_to_K.assign(AT.nnz, -1); GT_to_K.assign(GT.nnz, -1);
    ata = Gram(); gtg = Gram(); xx_P.clear(); xx_var.clear(); xx_ata.clear(); xx_gtg.clear();
    std::vector<int> diagK(nk, -1);
    if (mode == 0) {
        for (int j = 0; j < n; j++) {
            bool has_diag = false;
            for (int q = P.p[j]; q < P.p[j + 1]; q++) {
                if (P.i[q] > j) { error = "sparse_ldlt: P must be upper triangular"; return false; }
                P_to_K[q] = (int)Ki.size();
                if (P.i[q] == j) { has_diag = true; diagK[j] = (int)Ki.size(); }
                Ki.push_back(P.i[q]);
            }
            if (!has_diag) { diagK[j] = (int)Ki.size(); Ki.push_back(j); }
            Kp[j + 1] = (int)Ki.size();
        }
    } else {
        if (elim_eq) build_gram(AT, ata);
        if (elim_ineq) build_gram(GT, gtg);
        std::vector<int> slot(n, -1), rows;
        for (int j = 0; j < n; j++) {
            rows.clear();
            auto add = [&](int r) { if (slot[r] != j) { slot[r] = j; rows.push_back(r); } };
            for (int q = P.p[j]; q < P.p[j + 1]; q++) { if (P.i[q] > j) { error = "sparse_ldlt: P must be upper triangular"; return false; } add(P.i[q]); }
            add(j);
            if (elim_eq) for (int e = ata.colp[j]; e < ata.colp[j + 1]; e++) add(ata.rows[e]);
            if (elim_ineq) for (int e = gtg.colp[j]; e < gtg.colp[j + 1]; e++) add(gtg.rows[e]);
            std::sort(rows.begin(), rows.end());
            const int base = (int)Ki.size();
            for (size_t t = 0; t < rows.size(); t++) { Ki.push_back(rows[t]); xx_P.push_back(-1); xx_var.push_back(rows[t] == j ? j : -1); xx_ata.push_back(-1); xx_gtg.push_back(-1); }
            // positions: rows are sorted, binary search
            auto find = [&](int r) { return base + (int)(std::lower_bound(rows.begin(), rows.end(), r) - rows.begin()); };
            for (int q = P.p[j]; q < P.p[j + 1]; q++) { const int e = find(P.i[q]); P_to_K[q] = e; xx_P[e] = q; }
            diagK[j] = find(j);
            if (elim_eq) for (int e = ata.colp[j]; e < ata.colp[j + 1]; e++) xx_ata[find(ata.rows[e])] = e;
            if (elim_ineq) for (int e = gtg.colp[j]; e < gtg.colp[j + 1]; e++) xx_gtg[find(gtg.rows[e])] = e;
            Kp[j + 1] = (int)Ki.size();
        }
    }
    int col = n;
    if (!elim_eq) for (int c = 0; c < p; c++, col++) {
        for (int q = AT.p[c]; q < AT.p[c + 1]; q++) { AT_to_K[q] = (int)Ki.size(); Ki.push_back(AT.i[q]); }
        diagK[col] = (int)Ki.size(); Ki.push_back(col);
        Kp[col + 1] = (int)Ki.size();
    }
    if (!elim_ineq) for (int c = 0; c < m; c++, col++) {
        for (int q = GT.p[c]; q < GT.p[c + 1]; q++) { GT_to_K[q] = (int)Ki.size(); Ki.push_back(GT.i[q]); }
        diagK[col] = (int)Ki.size(); Ki.push_back(col);
        Kp[col + 1] = (int)Ki.size();
    }
    // ---- ordering
    if (user_perm) perm.assign(user_perm, user_perm + nk);
    else perm = minimum_degree_ordering(nk, Kp, Ki);
    lap("KKT pattern + ordering");
    const int nnzK = (int)Ki.size();
    std::vector<int> flag(nk, -1), Lnz(nk, 0);
    std::vector<std::pair<int, int>> extra;     // explicit structural zeros (row < col, permuted indices) added by supernode amalgamation
    nnzL_exact = -1.0; flops_exact = -1.0;
    bool have_mapped = false;
    std::vector<int> etree_m, Lnz_m;
    for (int pass = 0; pass < 3; pass++) {
    iperm.assign(nk, -1);
    for (int k = 0; k < nk; k++) { if (perm[k] < 0 || perm[k] >= nk || iperm[perm[k]] != -1) { error = "sparse_ldlt: invalid permutation"; return false; } iperm[perm[k]] = k; }
    // ---- permuted upper pattern with sorted rows + value map (utils.hpp:31-128)
    const int nnzPK = nnzK + (int)extra.size();
    std::vector<int> colcnt(nk + 1, 0), ecol(nnzK), erow(nnzK);
    for (int j = 0; j < nk; j++) for (int q = Kp[j]; q < Kp[j + 1]; q++) {
        const int a = iperm[Ki[q]], b = iperm[j];
        erow[q] = std::min(a, b); ecol[q] = std::max(a, b);
        colcnt[ecol[q] + 1]++;
    }
    for (const auto& x : extra) colcnt[x.second + 1]++;
    PKp.assign(nk + 1, 0);
    for (int j = 0; j < nk; j++) PKp[j + 1] = PKp[j] + colcnt[j + 1];
    std::vector<std::pair<int, int>> tmp(nnzPK);   // (row, K index or -1) grouped by column
    {
        std::vector<int> w(PKp.begin(), PKp.end() - 1);
        for (int q = 0; q < nnzK; q++) tmp[w[ecol[q]]++] = {erow[q], q};
        for (const auto& x : extra) tmp[w[x.second]++] = {x.first, -1};
    }
    PKi_rows.assign(nnzPK, 0); K_to_PK.assign(nnzK, 0);
    for (int j = 0; j < nk; j++) {
        std::sort(tmp.begin() + PKp[j], tmp.begin() + PKp[j + 1]);
        for (int t = PKp[j]; t < PKp[j + 1]; t++) { PKi_rows[t] = tmp[t].first; if (tmp[t].second >= 0) K_to_PK[tmp[t].second] = t; }
    }
    diagPK.assign(nk, -1);
    for (int v = 0; v < nk; v++) diagPK[v] = K_to_PK[diagK[v]];
    // ---- elimination tree and column counts, row by row (ldlt.hpp:42-99).  A postorder relabels the tree and keeps the counts,
    //      so the pass that follows the postordering takes them from the mapping instead of walking nnz(L) entries again.
    if (have_mapped) { etree.swap(etree_m); Lnz.swap(Lnz_m); have_mapped = false; }
    else if (!getenv("B200_SYMBOLIC_WALK")) etree_and_column_counts(nk, PKp, PKi_rows, etree, Lnz);
    else {
        etree.assign(nk, -1);
        std::fill(flag.begin(), flag.end(), -1); std::fill(Lnz.begin(), Lnz.end(), 0);
        for (int k = 0; k < nk; k++) {
            flag[k] = k;
            for (int q = PKp[k]; q < PKp[k + 1]; q++)
                for (int i = PKi_rows[q]; flag[i] != k; i = etree[i]) { if (etree[i] == -1) etree[i] = k; Lnz[i]++; flag[i] = k; }
        }
    }
    lap("permute + etree + column counts (one pass)");
    if (pass == 2) break;
    if (pass == 1) {
        // ---- relaxed supernode amalgamation.  Fundamental supernodes of these KKT systems are mostly single columns, so a
        // multifrontal sweep would pay its per-front overhead once per column.  A chain  s -> parent t  (t starts right after s)
        // is merged when padding s's columns to t's structure costs few explicit zeros; the padding is added to the PATTERN of
        // the permuted matrix (values stay 0), after which the symbolic phase below finds the merged supernodes by itself.
        if (getenv("B200_LDLT_NO_AMALG")) break;
        auto knob = [](const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; };
        const int k_abs = knob("B200_AMALG_ABS", 256), k_div = std::max(1, knob("B200_AMALG_DIV", 4)), k_grow = knob("B200_AMALG_GROW_PCT", 100);
        std::vector<int> fsp, fsof;
        std::vector<std::vector<int>> fU;
        if (!supernode_patterns(nk, etree, Lnz, PKp, PKi_rows, fsp, fsof, fU, error)) return false;
        lap("  amalgamation: supernode patterns");
        std::vector<int> pat;
        auto pattern_of = [&](int j) {                         // struct(L_j), sorted
            const int s3 = fsof[j], j1 = fsp[s3 + 1] - 1;
            pat.clear();
            for (int r = j + 1; r <= j1; r++) pat.push_back(r);
            pat.insert(pat.end(), fU[s3].begin(), fU[s3].end());
        };
        nnzL_exact = 0; for (int j = 0; j < nk; j++) nnzL_exact += Lnz[j];
        flops_exact = 0; for (int j = 0; j < nk; j++) { const double c = Lnz[j]; flops_exact += c * c + 2 * c; }
        std::vector<int> fs;                                   // first columns of the fundamental supernodes
        for (int j = 0; j < nk; j++) if (!(j > 0 && etree[j - 1] == j && Lnz[j - 1] == Lnz[j] + 1)) fs.push_back(j);
        fs.push_back(nk);
        const int nf = (int)fs.size() - 1;
        for (int s2 = 0; s2 < nf;) {
            const int a = fs[s2];
            int t2 = s2;                                       // group = fundamental supernodes s2..t2
            const int orig_first = Lnz[a];
            while (t2 + 1 < nf) {
                const int b = fs[t2 + 1] - 1;                  // last column of the group so far
                if (etree[b] != b + 1) break;
                const int bt = fs[t2 + 2] - 1;                 // last column of the candidate parent
                const int f_t = (bt - b) + Lnz[bt], newrows = f_t - Lnz[b], wcur = b - a + 1;
                const int padded_first = (bt - a) + Lnz[bt];
                const bool cheap = (long long)wcur * newrows <= k_abs || newrows <= std::max(2, Lnz[b] / k_div);
                if (!cheap || padded_first > orig_first + (int)((long long)orig_first * k_grow / 100) + 8) break;
                t2++;
            }
            if (t2 > s2) {
                const int bg = fs[t2 + 1] - 1;
                const std::vector<int>& Ug = fU[fsof[bg]];         // bg is the last column of its fundamental supernode: struct(L_bg) = U
                // target of column j = {j+1..bg} u Ug; struct(L_j) = {j+1..j1} u U(s3) for the fundamental supernode s3 = [.., j1] of j, so what
                // column j lacks is  M(s3) = ((j1, bg] u Ug) \ U(s3)  -- the same set for every column of s3: one merge per fundamental
                // supernode instead of one explicit pattern per column (config 3: 10 000 patterns of ~10 000 rows, 0.3 s)
                for (int s3 = s2; s3 <= t2; s3++) {
                    const int j1 = fs[s3 + 1] - 1;
                    if (j1 >= bg) break;                           // the group's last fundamental supernode already has the target structure
                    const std::vector<int>& Us = fU[fsof[j1]];
                    pat.clear();                                   // M(s3), sorted
                    const int* p0 = Us.data(); const int* p1 = Us.data() + Us.size();
                    for (int r = j1 + 1; r <= bg; r++) { while (p0 < p1 && *p0 < r) p0++; if (p0 == p1 || *p0 != r) pat.push_back(r); }
                    for (int u : Ug) { while (p0 < p1 && *p0 < u) p0++; if (p0 == p1 || *p0 != u) pat.push_back(u); }
                    if (pat.empty()) continue;
                    for (int j = std::max(fs[s3], a); j <= j1; j++) for (int r : pat) extra.push_back({j, r});
                }
            }
            s2 = t2 + 1;
        }
        lap("amalgamation");
        if (extra.empty()) break;
        continue;
    }
    // ---- postorder the elimination tree so that supernodes are runs of consecutive columns; an equivalent reordering: same
    //      fill, same arithmetic per column.  Children are visited by increasing column count (ties: by index), so the child
    //      with the largest structure -- the one a parent can form a supernode with, or be amalgamated with -- comes right
    //      before its parent.  A postordered input is left unchanged.
    std::vector<int> head(nk, -1), next(nk, -1), post; post.reserve(nk);
    {
        std::vector<int> byk(nk);
        for (int j = 0; j < nk; j++) byk[j] = j;
        std::stable_sort(byk.begin(), byk.end(), [&](int a, int b) { return Lnz[a] < Lnz[b]; });
        for (int t = nk - 1; t >= 0; t--) { const int j = byk[t]; if (etree[j] >= 0) { next[j] = head[etree[j]]; head[etree[j]] = j; } }
    }
    std::vector<int> stack;
    for (int r = 0; r < nk; r++) {
        if (etree[r] >= 0) continue;
        stack.push_back(r);
        while (!stack.empty()) {
            const int v = stack.back();
            if (head[v] >= 0) { const int c = head[v]; head[v] = next[c]; stack.push_back(c); }
            else { post.push_back(v); stack.pop_back(); }
        }
    }
    bool identity = true;
    for (int k = 0; k < nk; k++) if (post[k] != k) { identity = false; break; }
    if (identity) continue;
    std::vector<int> np(nk), ipost(nk);
    for (int k = 0; k < nk; k++) { np[k] = perm[post[k]]; ipost[post[k]] = k; }
    perm.swap(np);
    etree_m.assign(nk, -1); Lnz_m.assign(nk, 0);
    for (int k = 0; k < nk; k++) { const int e = etree[post[k]]; etree_m[k] = e >= 0 ? ipost[e] : -1; Lnz_m[k] = Lnz[post[k]]; }
    have_mapped = true;
    }
    const int nnzPK = (int)PKi_rows.size();
    Lp.assign(nk + 1, 0);
    for (int k = 0; k < nk; k++) Lp[k + 1] = Lp[k] + Lnz[k];
    Li.assign(Lp[nk], 0);
    // Pattern of L, supernode by supernode: the columns of a supernode (a chain j0 -> ... -> j1 of the elimination tree whose
    // column counts drop by one) have the patterns {j+1..j1} u U, U = struct(L_j1), and
    // U = rows > j1 of [ the matrix columns of the supernode  u  the U's of its child supernodes ].
    // One merge per supernode and sequential writes instead of nnz(L) scattered writes along elimination-tree walks
    // (config 3: 1.3 s -> 0.2 s).  The row view is built on demand (build_row_view).
    {
        std::vector<int> sp, sof;
        std::vector<std::vector<int>> U;
        if (!supernode_patterns(nk, etree, Lnz, PKp, PKi_rows, sp, sof, U, error)) return false;
        for (int s2 = 0; s2 + 1 < (int)sp.size(); s2++) {
            const int j0 = sp[s2], j1 = sp[s2 + 1] - 1;
            for (int j = j0; j <= j1; j++) {
                int pos = Lp[j];
                for (int r = j + 1; r <= j1; r++) Li[pos++] = r;
                for (int r : U[s2]) Li[pos++] = r;
                if (pos != Lp[j + 1]) { error = "sparse_ldlt: internal error (supernodal pattern does not match the column counts)"; return false; }
            }
        }
    }
    Rp.clear(); Rcol.clear(); Rpos.clear();
    lap("pattern of L + row view");
    // ---- scatter map of the permuted matrix into L / D
    PK_to_L.assign(want_level_maps ? nnzPK : 0, 0);
    if (want_level_maps) for (int j = 0; j < nk; j++) for (int q = PKp[j]; q < PKp[j + 1]; q++) {
        const int i = PKi_rows[q];
        if (i == j) { PK_to_L[q] = -(j + 1); continue; }
        const int* b = &Li[Lp[i]]; const int* e = &Li[Lp[i + 1]];
        const int* it = std::lower_bound(b, e, j);
        if (it == e || *it != j) { error = "sparse_ldlt: internal error (entry of A missing in L)"; return false; }
        PK_to_L[q] = (int)(it - &Li[0]);
    }
    lap("PK_to_L");
    // ---- level sets of the elimination tree
    level.assign(nk, 0);
    int maxl = 0;
    for (int j = 0; j < nk; j++) { if (etree[j] >= 0) level[etree[j]] = std::max(level[etree[j]], level[j] + 1); maxl = std::max(maxl, level[j]); }
    level_ptr.assign(maxl + 2, 0);
    for (int j = 0; j < nk; j++) level_ptr[level[j] + 1]++;
    for (int l = 0; l <= maxl; l++) level_ptr[l + 1] += level_ptr[l];
    level_cols.assign(nk, 0);
    { std::vector<int> w(level_ptr.begin(), level_ptr.end() - 1); for (int j = 0; j < nk; j++) level_cols[w[level[j]]++] = j; }
    // ---- supernodes (maximal runs j, j+1 with parent(j) = j+1 and |struct(L_j)| = |struct(L_{j+1})| + 1) and the multifrontal schedule
    sup_ptr.clear();
    std::vector<int> sup_of(nk, 0);
    for (int j = 0; j < nk; j++) {
        const bool cont = j > 0 && etree[j - 1] == j && (Lp[j] - Lp[j - 1]) == (Lp[j + 1] - Lp[j]) + 1;
        if (!cont) sup_ptr.push_back(j);
        sup_of[j] = (int)sup_ptr.size() - 1;
    }
    nsup = (int)sup_ptr.size();
    sup_ptr.push_back(nk);
    std::vector<int> psup(nsup, -1);
    child_ptr.assign(nsup + 1, 0);
    for (int s2 = 0; s2 < nsup; s2++) { const int j1 = sup_ptr[s2 + 1] - 1; if (etree[j1] >= 0) { psup[s2] = sup_of[etree[j1]]; child_ptr[psup[s2] + 1]++; } }
    for (int s2 = 0; s2 < nsup; s2++) child_ptr[s2 + 1] += child_ptr[s2];
    child_idx.assign(child_ptr[nsup], 0);
    { std::vector<int> w(child_ptr.begin(), child_ptr.end() - 1); for (int s2 = 0; s2 < nsup; s2++) if (psup[s2] >= 0) child_idx[w[psup[s2]]++] = s2; }
    auto front_row = [&](int s2, int row) -> int {          // position of global row `row` in the front of supernode s2, or -1
        const int j0 = sup_ptr[s2], j1 = sup_ptr[s2 + 1] - 1;
        if (row >= j0 && row <= j1) return row - j0;
        const int* b = &Li[0] + Lp[j1]; const int* e = &Li[0] + Lp[j1 + 1];
        const int* it = std::lower_bound(b, e, row);
        if (it == e || *it != row) return -1;
        return (j1 - j0 + 1) + (int)(it - b);
    };
    fmax = 0;
    rel_ptr.assign(nsup + 1, 0);
    for (int s2 = 0; s2 < nsup; s2++) {
        const int j1 = sup_ptr[s2 + 1] - 1, us = Lp[j1 + 1] - Lp[j1];
        rel_ptr[s2 + 1] = rel_ptr[s2] + us;
        fmax = std::max(fmax, (j1 - sup_ptr[s2] + 1) + us);
    }
    rel_idx.assign(rel_ptr[nsup], 0);
    for (int s2 = 0; s2 < nsup; s2++) {
        const int j1 = sup_ptr[s2 + 1] - 1;
        for (int t = 0; t < rel_ptr[s2 + 1] - rel_ptr[s2]; t++) {
            const int r = front_row(psup[s2], Li[Lp[j1] + t]);      // us > 0 implies a parent
            if (r < 0) { error = "sparse_ldlt: internal error (update row missing in the parent front)"; return false; }
            rel_idx[rel_ptr[s2] + t] = r;
        }
    }
    asm_ptr.assign(nsup + 1, 0);
    for (int j = 0; j < nk; j++) for (int q = PKp[j]; q < PKp[j + 1]; q++) asm_ptr[sup_of[PKi_rows[q]] + 1]++;
    for (int s2 = 0; s2 < nsup; s2++) asm_ptr[s2 + 1] += asm_ptr[s2];
    asm_q.assign(nnzPK, 0); asm_pos.assign(nnzPK, 0);
    {
        std::vector<int> w(asm_ptr.begin(), asm_ptr.end() - 1);
        for (int j = 0; j < nk; j++) for (int q = PKp[j]; q < PKp[j + 1]; q++) {
            const int i = PKi_rows[q], s2 = sup_of[i];          // upper entry (i <= j) = lower entry (row j, column i) of the front of i's supernode
            const int j1 = sup_ptr[s2 + 1] - 1, f = (j1 - sup_ptr[s2] + 1) + (Lp[j1 + 1] - Lp[j1]);
            const int r = front_row(s2, j);
            if (r < 0) { error = "sparse_ldlt: internal error (matrix entry outside its front)"; return false; }
            const int t = w[s2]++;
            asm_q[t] = q; asm_pos[t] = r + (i - sup_ptr[s2]) * f;
        }
    }
    lap("levels, supernodes, relative indices, assembly map");
    // update-matrix stack: children sit on top of the stack when their parent is assembled (postorder)
    upd_off.assign(nsup, 0);
    long long top = 0; upd_total = 0;
    for (int s2 = 0; s2 < nsup; s2++) {
        for (int c = child_ptr[s2]; c < child_ptr[s2 + 1]; c++) { const long long uc = rel_ptr[child_idx[c] + 1] - rel_ptr[child_idx[c]]; top -= uc * uc; }
        const long long us = rel_ptr[s2 + 1] - rel_ptr[s2];
        upd_off[s2] = top; top += us * us;
        upd_total = std::max(upd_total, top);
    }
    if (getenv("B200_DEBUG_SYMBOLIC")) {
        double sf2 = 0, su2 = 0, piv_work = 0; int big = 0, w1 = 0;
        for (int s2 = 0; s2 < nsup; s2++) {
            const double ws = sup_ptr[s2 + 1] - sup_ptr[s2], us = rel_ptr[s2 + 1] - rel_ptr[s2], f = ws + us;
            sf2 += f * f; su2 += us * us; big += f > 150; w1 += ws == 1;
            for (int k = 0; k < (int)ws; k++) piv_work += (f - k) * (f - k) / 2;
        }
        for (int s2 = 0; s2 < nsup; s2++) { const int ws = sup_ptr[s2 + 1] - sup_ptr[s2], us = rel_ptr[s2 + 1] - rel_ptr[s2];
            if (ws + us > 150) fprintf(stderr, "  big front: s=%d j0=%d ws=%d us=%d parent=%d nchild=%d\n", s2, sup_ptr[s2], ws, us, psup[s2], child_ptr[s2 + 1] - child_ptr[s2]); }
        fprintf(stderr, "[sparse_ldlt symbolic] nk=%d nnzL=%zu nsup=%d (width 1: %d) fmax=%d fronts>150: %d  sum f^2=%.3g  sum us^2=%.3g  pivot-update entries=%.3g  flops=%.3g upd_total=%lld\n",
                nk, Li.size(), nsup, w1, fmax, big, sf2, su2, piv_work, factor_flops(), upd_total);
    }
    return true;
}
// Row view of L by one counting transpose (columns of a row arrive in increasing order: no sorting).  skip_sup (optional, per
// supernode): entries whose row AND column lie in the same flagged supernode are left out -- the blocked dense solves of the
// whole-GPU schedule read those triangles from the column storage, and for a 10 000-column root they are 98 % of nnz(L).
void LdltSymbolic::build_row_view(const std::vector<char>* skip_sup) {
    std::vector<int> sup_of;
    if (skip_sup) { sup_of.assign(nk, 0); for (int s2 = 0; s2 < nsup; s2++) for (int j = sup_ptr[s2]; j < sup_ptr[s2 + 1]; j++) sup_of[j] = s2; }
    auto skipped = [&](int i, int k) { return skip_sup && sup_of[i] == sup_of[k] && (*skip_sup)[sup_of[i]]; };
    Rp.assign(nk + 1, 0);
    for (int i = 0; i < nk; i++) for (int pos = Lp[i]; pos < Lp[i + 1]; pos++) if (!skipped(i, Li[pos])) Rp[Li[pos] + 1]++;
    for (int k = 0; k < nk; k++) Rp[k + 1] += Rp[k];
    Rcol.assign(Rp[nk], 0); Rpos.assign(Rp[nk], 0);
    std::vector<int> rfill(Rp.begin(), Rp.end() - 1);
    for (int i = 0; i < nk; i++) for (int pos = Lp[i]; pos < Lp[i + 1]; pos++) {
        const int k = Li[pos];
        if (skipped(i, k)) continue;
        const int t = rfill[k]++; Rcol[t] = i; Rpos[t] = pos;
    }
}
double LdltSymbolic::factor_flops() const {
    if (flops_exact >= 0) return flops_exact;      // algorithmic figure: explicit zeros of amalgamated supernodes are not counted
    double f = 0;
    for (int j = 0; j < nk; j++) { const double c = Lp[j + 1] - Lp[j]; f += c * c + 2 * c; }
    return f;
}

// =====================================================================================================
// device kernels
// =====================================================================================================
__global__ void ldlt_scatter_kernel(const int* map, int nnz, int nnzPK, const double* vals, double* PKx) {
    const int b = blockIdx.y;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < nnz) PKx[(size_t)b * nnzPK + map[q]] = vals[(size_t)b * nnz + q];
}
// diagonal of the KKT matrix for this iteration (kkt_full.hpp:172-210).  pk / mk: sizes of the y / z blocks kept in the KKT;
// skip_x: the condensed modes rebuild the whole top-left block in ldlt_cond_assemble_kernel instead.
__global__ void ldlt_set_diag_kernel(const int* diagPK, int n, int pk, int mk, int skip_x, int nnzPK, const double* P_diag, const double* x_reg, const double* delta,
                                     const double* z_reg, double* PKx, const int* active) {
    const int b = blockIdx.y;
    if (active && !active[b]) return;
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    const int nk = n + pk + mk;
    if (v >= nk) return;
    double val;
    if (v < n) { if (skip_x) return; val = P_diag[(size_t)b * n + v] + x_reg[(size_t)b * n + v]; }
    else if (v < n + pk) val = -delta[b];
    else val = -z_reg[(size_t)b * mk + (v - n - pk)];
    PKx[(size_t)b * nnzPK + diagPK[v]] = val;
}
// values of upper(A^T A) from the contribution lists (update_AT_A, kkt_eq_eliminated.hpp:223-245): one thread per entry,
// summands in ascending k like the reference's Gustavson loop
__global__ void ldlt_gram_values_kernel(const int* ptr, const int* pa, const int* pb, int nent, int nnzM, const double* vals, double* out) {
    const int b = blockIdx.y;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nent) return;
    const double* v = vals + (size_t)b * nnzM;
    double acc = 0.0;
    for (int t = ptr[e]; t < ptr[e + 1]; t++) acc += v[pb[t]] * v[pa[t]];
    out[(size_t)b * nent + e] = acc;
}
// top-left block of the condensed KKT:  0 + P + x_reg (diagonal) + delta^-1 (A^T A) + G^T Z^-1 G, in the reference's order of
// additions (kkt_all_eliminated.hpp:108-160); the G^T Z^-1 G entry is summed as (G_kj * G_ki) / z_reg_k, k ascending (:199-220)
__global__ void ldlt_cond_assemble_kernel(int nxx, const int* xx_P, const int* xx_var, const int* xx_ata, const int* xx_gtg, const int* xx_target,
                                          const int* gptr, const int* gpa, const int* gpb, const int* gcolof, int n, int m, int nnzP, int nnzG, int nata, int nnzPK,
                                          const double* Px, const double* GTx, const double* AtA, const double* x_reg, const double* delta, const double* z_reg,
                                          double* PKx, const int* active) {
    const int b = blockIdx.y;
    if (active && !active[b]) return;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nxx) return;
    double v = 0.0;
    const int ps = xx_P[e], j = xx_var[e], a = xx_ata[e], g = xx_gtg[e];
    if (ps >= 0) v += Px[(size_t)b * nnzP + ps];
    if (j >= 0) v += x_reg[(size_t)b * n + j];
    if (a >= 0) v += (1.0 / delta[b]) * AtA[(size_t)b * nata + a];
    if (g >= 0) {
        const double* G = GTx + (size_t)b * nnzG;
        const double* zr = z_reg + (size_t)b * m;
        double acc = 0.0;
        for (int t = gptr[g]; t < gptr[g + 1]; t++) { const int ia = gpa[t]; acc += G[gpb[t]] * G[ia] / zr[gcolof[ia]]; }
        v += acc;
    }
    PKx[(size_t)b * nnzPK + xx_target[e]] = v;
}
__global__ void ldlt_store_scalings_kernel(int m, const double* delta, const double* z_reg, double* dlt, double* zinv, const int* active) {
    const int b = blockIdx.y;
    if (active && !active[b]) return;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k == 0) dlt[b] = delta[b];
    if (k < m) zinv[(size_t)b * m + k] = 1.0 / z_reg[(size_t)b * m + k];
}
__global__ void ldlt_zero_kernel(double* Lx, size_t nnzL, const int* active) {
    const int b = blockIdx.y;
    if (active && !active[b]) return;
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < nnzL) Lx[(size_t)b * nnzL + e] = 0.0;
}
__global__ void ldlt_init_kernel(const int* PK_to_L, int nnzPK, size_t nnzL, int nk, const double* PKx, double* Lx, double* Dv, const int* active) {
    const int b = blockIdx.y;
    if (active && !active[b]) return;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nnzPK) return;
    const int t = PK_to_L[q];
    const double v = PKx[(size_t)b * nnzPK + q];
    if (t >= 0) Lx[(size_t)b * nnzL + t] = v; else Dv[(size_t)b * nk + (-t - 1)] = v;
}
// one warp per (column of this level, instance): left-looking update + scaling
__global__ void ldlt_level_kernel(const int* cols, int ncols, const int* Lp, const int* Li, const int* Rp, const int* Rcol, const int* Rpos,
                                  size_t nnzL, int nk, double* Lx_all, double* Dv_all, double* Dinv_all, int* fail, const int* active) {
    const int b = blockIdx.y;
    if (active && !active[b]) return;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= ncols) return;
    const int j = cols[warp];
    double* Lx = Lx_all + (size_t)b * nnzL;
    double* Dv = Dv_all + (size_t)b * nk;
    const int cj0 = Lp[j], cj1 = Lp[j + 1];
    double d = Dv[j];
    for (int t = Rp[j]; t < Rp[j + 1]; t++) {
        const int k = Rcol[t], pos = Rpos[t];
        const double ljk = Lx[pos];
        const double w = ljk * Dv[k];
        d -= ljk * w;
        for (int e = pos + 1 + lane; e < Lp[k + 1]; e += 32) {
            const int r = Li[e];
            int lo = cj0, hi = cj1 - 1;           // r is guaranteed to be in column j's pattern
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (Li[mid] < r) lo = mid + 1; else hi = mid; }
            Lx[lo] -= Lx[e] * w;
        }
        __syncwarp();
    }
    if (lane == 0) {
        if (d == 0.0 && fail[b] == 0) fail[b] = j + 1;      // ldlt.hpp:161
        Dv[j] = d;
        Dinv_all[(size_t)b * nk + j] = 1.0 / d;
    }
    for (int e = cj0 + lane; e < cj1; e += 32) Lx[e] /= d;
}
// rhs (x | y | z blocks) -> permuted work vector, and back (ordering.hpp:102-124, sparse/kkt.hpp:113-147)
__global__ void ldlt_gather_rhs_kernel(const int* perm, int n, int p, int m, const double* rx, const double* ry, const double* rz, double* work, const int* active) {
    const int b = blockIdx.y;
    if (active && !active[b]) return;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int nk = n + p + m;
    if (j >= nk) return;
    const int v = perm[j];
    double val;
    if (v < n) val = rx[(size_t)b * n + v]; else if (v < n + p) val = ry[(size_t)b * p + (v - n)]; else val = rz[(size_t)b * m + (v - n - p)];
    work[(size_t)b * nk + j] = val;
}
__global__ void ldlt_scatter_lhs_kernel(const int* perm, int n, int p, int m, const double* work, double* lx, double* ly, double* lz, const int* active) {
    const int b = blockIdx.y;
    if (active && !active[b]) return;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int nk = n + p + m;
    if (j >= nk) return;
    const int v = perm[j];
    const double val = work[(size_t)b * nk + j];
    if (v < n) lx[(size_t)b * n + v] = val; else if (v < n + p) ly[(size_t)b * p + (v - n)] = val; else lz[(size_t)b * m + (v - n - p)] = val;
}
__global__ void ldlt_fwd_level_kernel(const int* cols, int ncols, const int* Rp, const int* Rcol, const int* Rpos, size_t nnzL, int nk, const double* Lx_all,
                                      double* work_all, const int* active) {
    const int b = blockIdx.y;
    if (active && !active[b]) return;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ncols) return;
    const int j = cols[t];
    const double* Lx = Lx_all + (size_t)b * nnzL;
    double* w = work_all + (size_t)b * nk;
    double acc = w[j];
    for (int q = Rp[j]; q < Rp[j + 1]; q++) acc -= Lx[Rpos[q]] * w[Rcol[q]];
    w[j] = acc;
}
__global__ void ldlt_dscale_kernel(int nk, const double* Dinv, double* work, const int* active) {
    const int b = blockIdx.y;
    if (active && !active[b]) return;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < nk) work[(size_t)b * nk + j] *= Dinv[(size_t)b * nk + j];
}
__global__ void ldlt_bwd_level_kernel(const int* cols, int ncols, const int* Lp, const int* Li, size_t nnzL, int nk, const double* Lx_all, double* work_all,
                                      const int* active) {
    const int b = blockIdx.y;
    if (active && !active[b]) return;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ncols) return;
    const int j = cols[t];
    const double* Lx = Lx_all + (size_t)b * nnzL;
    double* w = work_all + (size_t)b * nk;
    double acc = w[j];
    for (int e = Lp[j]; e < Lp[j + 1]; e++) acc -= Lx[e] * w[Li[e]];
    w[j] = acc;
}
__global__ void ldlt_clear_fail_kernel(int* fail, const int* active, int batch) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < batch && (!active || active[b])) fail[b] = 0;
}
__global__ void ldlt_fail_to_ok_kernel(const int* fail, const int* active, int* ok, int batch) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < batch && (!active || active[b])) ok[b] = fail[b] ? 0 : 1;
}

// =====================================================================================================
static void upload(DevBuf<int>& d, const std::vector<int>& h) {
    d.alloc(std::max<size_t>(h.size(), 1));
    if (!h.empty()) B200_CUDA(cudaMemcpy(d.get(), h.data(), h.size() * sizeof(int), cudaMemcpyHostToDevice));
}
static MfDev make_mf(const SparseLdltBatchedKKT& K) {
    MfDev M;
    M.hdr = K.d_hdr.get(); M.crec = K.d_crec.get(); M.rel_idx = K.d_rel_idx.get(); M.asm_pos = K.d_asm_pos.get(); M.Li = K.d_Li.get(); M.perm = K.d_perm.get();
    M.upd_off = K.d_upd_off.get();
    M.nsup = K.S.nsup; M.nk = K.S.nk; M.n = K.n; M.p = K.S.pk; M.m = K.S.mk; M.front_smem_rows = K.front_smem_rows; M.fmax = K.S.fmax;
    M.upd_total = std::max<long long>(K.S.upd_total, 1);
    M.nnzL = std::max<size_t>(K.S.Li.size(), 1); M.nnzPK = K.S.PKi_rows.size();
    M.prof = K.d_prof.n ? K.d_prof.get() : nullptr;
    { const char* e = getenv("B200_MF_BIG"); M.big_right_looking = (e && std::string(e) == "right") ? 1 : 0; }
    return M;
}

// =====================================================================================================
// whole-GPU schedule (sparse_wide.cuh)
// =====================================================================================================
SparseLdltBatchedKKT::~SparseLdltBatchedKKT() {
    if (d_prof.n) {
        long long h[8]; device_synchronize_shared();
        if (cudaMemcpy(h, d_prof.get(), sizeof h, cudaMemcpyDeviceToHost) == cudaSuccess)
            fprintf(stderr, "[mf_factor_kernel phase clocks, CTA 0] zero+scatter %lld  extend-add %lld  eliminate(smem) %lld  eliminate(HBM) %lld  schur store %lld\n", h[0], h[1], h[2], h[3], h[4]);
    }
    if (ev_col) cudaEventDestroy(ev_col);
    if (ev_panel) cudaEventDestroy(ev_panel);
    if (aux_stream) cudaStreamDestroy(aux_stream);
    if (ev_p2) cudaEventDestroy(ev_p2);
    if (ev_r2) cudaEventDestroy(ev_r2);
    if (aux2_stream) cudaStreamDestroy(aux2_stream);
}

void SparseLdltBatchedKKT::build_wide() {
    if (!aux_stream && !getenv("B200_WIDE_NO_LOOKAHEAD")) {
        int prio_lo = 0, prio_hi = 0;
        B200_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        B200_CUDA(cudaStreamCreateWithPriority(&aux_stream, cudaStreamNonBlocking, prio_hi));      // panel CTAs go first when SMs free up
        B200_CUDA(cudaEventCreateWithFlags(&ev_col, cudaEventDisableTiming));
        B200_CUDA(cudaEventCreateWithFlags(&ev_panel, cudaEventDisableTiming));
        if (!getenv("B200_WIDE_NO_WINDOW_SPLIT")) {
            B200_CUDA(cudaStreamCreateWithPriority(&aux2_stream, cudaStreamNonBlocking, prio_hi));
            B200_CUDA(cudaEventCreateWithFlags(&ev_p2, cudaEventDisableTiming));
            B200_CUDA(cudaEventCreateWithFlags(&ev_r2, cudaEventDisableTiming));
        }
    }
    const int nsup = S.nsup, nk = S.nk;
    auto knob = [](const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; };
    wide_group = std::min(8, std::max(1, knob("B200_WIDE_GROUP", 4)));

    wide_sb = 128;
    { const int v = knob("B200_WIDE_SB", 128); for (int c : {8, 16, 32, 64, 128}) if (v == c) wide_sb = c; }     // power of two <= 128
    const int ws_min = std::max(2, knob("B200_WIDE_WS", 64));          // supernodes at least this wide are solved blocked over the GPU
    std::vector<int> sup_of(nk, 0), psup(nsup, -1), slevel(nsup, 0);
    for (int s2 = 0; s2 < nsup; s2++) for (int j = S.sup_ptr[s2]; j < S.sup_ptr[s2 + 1]; j++) sup_of[j] = s2;
    int maxl = 0;
    for (int s2 = 0; s2 < nsup; s2++) {          // postorder: children come before their parent
        const int j1 = S.sup_ptr[s2 + 1] - 1;
        if (S.etree[j1] >= 0) { psup[s2] = sup_of[S.etree[j1]]; slevel[psup[s2]] = std::max(slevel[psup[s2]], slevel[s2] + 1); }
        maxl = std::max(maxl, slevel[s2]);
    }
    // every supernode keeps its own update slot (siblings of different subtrees are in flight at the same time)
    std::vector<long long> off(nsup, 0);
    upd_total_w = 0;
    for (int s2 = 0; s2 < nsup; s2++) { const long long us = S.rel_ptr[s2 + 1] - S.rel_ptr[s2]; off[s2] = upd_total_w; upd_total_w += us * us; }
    std::vector<int> crecw((size_t)4 * std::max<size_t>(S.child_idx.size(), 1), 0);
    for (size_t c = 0; c < S.child_idx.size(); c++) {
        const int cs = S.child_idx[c];
        crecw[4 * c] = S.rel_ptr[cs + 1] - S.rel_ptr[cs]; crecw[4 * c + 1] = S.rel_ptr[cs];
        crecw[4 * c + 2] = (int)(unsigned)(off[cs] & 0xffffffffll); crecw[4 * c + 3] = (int)(off[cs] >> 32);
    }
    upload(d_crecw, crecw);
    d_upd_off_w.alloc(std::max<size_t>(nsup, 1));
    B200_CUDA(cudaMemcpy(d_upd_off_w.get(), off.data(), (size_t)nsup * sizeof(long long), cudaMemcpyHostToDevice));
    upd.alloc((size_t)batch * (size_t)std::max<long long>(upd_total_w, 1));
    // ---- steps by level
    std::vector<std::vector<int>> by_level(maxl + 1);
    for (int s2 = 0; s2 < nsup; s2++) by_level[slevel[s2]].push_back(s2);
    std::vector<int> list_f, list_s, pull_ptr, pull_child, pull_cc;
    wf_steps.clear(); ws_steps.clear(); wfronts.clear(); wf_smem.clear();
    front_stride = 0;
    size_t small_smem_max = 0;
    long long tinv_blocks = 0;
    for (int l = 0; l <= maxl; l++) {
        const int fb = (int)list_f.size(), sb0 = (int)list_s.size();
        int fmax_l = 0;
        for (int s2 : by_level[l]) {
            const int ws = S.sup_ptr[s2 + 1] - S.sup_ptr[s2], us = S.rel_ptr[s2 + 1] - S.rel_ptr[s2], f = ws + us;
            if (f <= front_smem_rows) { list_f.push_back(s2); fmax_l = std::max(fmax_l, f); }
            if (ws < ws_min) list_s.push_back(s2);
        }
        if ((int)list_f.size() > fb) {
            const int fpad = (fmax_l + 1) & ~1;
            const size_t smem = sizeof(double) * ((size_t)((fpad + fpad / 2 + 3) & ~3) + (size_t)fmax_l * fmax_l);
            wf_steps.push_back({0, fb, (int)list_f.size() - fb, fpad});
            wf_smem.push_back(smem); small_smem_max = std::max(small_smem_max, smem);
        }
        if ((int)list_s.size() > sb0) ws_steps.push_back({0, sb0, (int)list_s.size() - sb0, 0});
        for (int s2 : by_level[l]) {
            const int j0 = S.sup_ptr[s2], j1 = S.sup_ptr[s2 + 1] - 1, ws = j1 - j0 + 1, us = S.rel_ptr[s2 + 1] - S.rel_ptr[s2], f = ws + us;
            if (f > front_smem_rows) {
                if (us > 65535) throw std::runtime_error("sparse_ldlt: update matrix too large for the wide schedule");
                WFront w{};
                w.s = s2; w.j0 = j0; w.ws = ws; w.us = us; w.f = f; w.shift = ws & 1; w.ld = round_up(f + w.shift, 8); w.lp0 = S.Lp[j0];
                w.ab = S.asm_ptr[s2]; w.an = S.asm_ptr[s2 + 1] - S.asm_ptr[s2]; w.off = off[s2];
                w.nchild = S.child_ptr[s2 + 1] - S.child_ptr[s2];
                w.pull_begin = (int)pull_ptr.size();
                // pull lists: for every front column, the (child record, child column) pairs that map onto it, children in order
                std::vector<int> cnt(f + 1, 0);
                for (int c = S.child_ptr[s2]; c < S.child_ptr[s2 + 1]; c++) { const int cs = S.child_idx[c];
                    for (int t = S.rel_ptr[cs]; t < S.rel_ptr[cs + 1]; t++) cnt[S.rel_idx[t] + 1]++; }
                const int base = (int)pull_child.size();
                for (int c = 0; c < f; c++) cnt[c + 1] += cnt[c];
                for (int c = 0; c <= f; c++) pull_ptr.push_back(base + cnt[c]);
                pull_child.resize(base + cnt[f]); pull_cc.resize(base + cnt[f]);
                std::vector<int> fillp(cnt.begin(), cnt.end() - 1);
                for (int c = S.child_ptr[s2]; c < S.child_ptr[s2 + 1]; c++) { const int cs = S.child_idx[c];
                    for (int t = S.rel_ptr[cs]; t < S.rel_ptr[cs + 1]; t++) { const int q = base + fillp[S.rel_idx[t]]++; pull_child[q] = c; pull_cc[q] = t - S.rel_ptr[cs]; } }
                front_stride = std::max(front_stride, (long long)w.ld * f);
                wf_steps.push_back({1, (int)wfronts.size(), 0, 0});
                wfronts.push_back(w);
            }
            if (ws >= ws_min) { ws_steps.push_back({1, s2, (int)tinv_blocks, 0}); tinv_blocks += ceil_div(ws, wide_sb); }
        }
    }
    {   // row view for the pull-form forward solves, without the triangles of the supernodes that are solved blocked
        std::vector<char> skip(nsup, 0);
        for (int s2 = 0; s2 < nsup; s2++) skip[s2] = (S.sup_ptr[s2 + 1] - S.sup_ptr[s2]) >= ws_min;
        S.build_row_view(&skip);
        upload(d_Rp, S.Rp); upload(d_Rcol, S.Rcol); upload(d_Rpos, S.Rpos);
    }
    upload(d_wlist_f, list_f); upload(d_wlist_s, list_s); upload(d_pull_ptr, pull_ptr); upload(d_pull_child, pull_child); upload(d_pull_cc, pull_cc);
    if (front_stride > 0) bigfront.alloc((size_t)batch * (size_t)front_stride);
    wtmp.alloc((size_t)batch * wide_sb); wcounter.alloc(batch);
    tinv_stride = tinv_blocks * wide_sb * wide_sb;
    Tcm.alloc(std::max<size_t>((size_t)batch * (size_t)tinv_stride, 1)); Trm.alloc(std::max<size_t>((size_t)batch * (size_t)tinv_stride, 1));
    B200_CUDA(cudaMemset(wcounter.get(), 0, sizeof(unsigned) * batch));
    const size_t sbs = (size_t)wide_sb;
    allow_dynamic_smem(mfw_small_kernel, (size_t)((int)std::max<size_t>(small_smem_max, 48 * 1024)));
    allow_dynamic_smem(mfw_block_inverse_kernel, (size_t)((int)std::max<size_t>(sizeof(double) * sbs * (sbs + 1), 48 * 1024)));
    allow_dynamic_smem(mfw_panel_kernel, (size_t)((int)MW_PANEL_SMEM));
    { const char* e = getenv("B200_WIDE_GRAPH"); wide_graph = e ? e[0] == '1' : wfronts.size() > 4; }
    if (getenv("B200_DEBUG_SYMBOLIC"))
        fprintf(stderr, "[sparse_ldlt wide] levels=%d factor steps=%zu (HBM fronts %zu) solve steps=%zu upd_total=%lld front_stride=%lld\n", maxl + 1, wf_steps.size(),
                wfronts.size(), ws_steps.size(), upd_total_w, front_stride);
}

void SparseLdltBatchedKKT::factor_wide(const int* active) {
    const MfDev M = make_mf(*this);
    const size_t nnzL = std::max<size_t>(S.Li.size(), 1);
    const int nk = S.nk;
    B200_LAUNCH(ldlt_clear_fail_kernel, ceil_div(batch, 256), 256, 0, stream, fail.get(), active, batch);
    size_t small_i = 0;
    for (const WStep& st : wf_steps) {
        if (st.kind == 0) {
            dim3 g(st.b, batch);
            B200_LAUNCH(mfw_small_kernel, g, MF_T, wf_smem[small_i], stream, M, d_wlist_f.get() + st.a, st.c, d_crecw.get(), d_upd_off_w.get(), PKx.get(), Lx.get(), Dv.get(),
                        Dinv.get(), upd.get(), upd_total_w, fail.get(), active);
            small_i++;
            continue;
        }
        const WFront& w = wfronts[st.a];
        double* F = bigfront.get();
        const long long count = (long long)w.ld * w.f;
        { dim3 g((unsigned)std::min<long long>((count / 2 + 255) / 256, 148 * 16), batch); B200_LAUNCH(mfw_zero_kernel, g, 256, 0, stream, F, front_stride, count, active); }
        if (w.an > 0) { dim3 g(ceil_div(w.an, 256), batch);
            B200_LAUNCH(mfw_scatter_kernel, g, 256, 0, stream, F, front_stride, w.ld, w.shift, w.f, d_asm_pos.get(), PKx.get(), M.nnzPK, w.ab, w.an, active); }
        if (w.nchild > 0) { dim3 g(ceil_div(w.f, MW_T / 32), batch);
            B200_LAUNCH(mfw_pull_kernel, g, MW_T, 0, stream, F, front_stride, w.ld, w.shift, w.f, d_pull_ptr.get() + w.pull_begin, d_pull_child.get(), d_pull_cc.get(),
                        d_crecw.get(), d_rel_idx.get(), upd.get(), upd_total_w, active); }
        // Blocked right-looking LDL^T, two levels, with look-ahead.  Panels of MW_NB columns are taken in GROUPS of `wide_group`:
        // inside a group, panel j is followed by a narrow WINDOW update of the group's remaining columns only (K = 64); the rest
        // of the front gets one FAR update per group with K = 64 * wide_group (the DMMA tile kernel needs a deep contraction to
        // amortise its epilogue: 13 TFLOP/s at K = 64, 21+ at K = 256).  Look-ahead: the first tile columns of a far update cover
        // the next group's columns, so the next group's panel / window chain runs on the (high-priority) auxiliary stream while
        // the remaining tile columns of the far update run on the main stream -- they touch disjoint columns of F.
        auto panel = [&](int pk0, int pnb, cudaStream_t st) {
            const int R = w.f - pk0 - pnb;
            dim3 g(std::max(1, ceil_div(R, MW_T)), batch);
            B200_LAUNCH(mfw_panel_kernel, g, MW_T, MW_PANEL_SMEM, st, F, front_stride, w.ld, w.shift, w.f, pk0, pnb, w.j0, w.lp0, Lx.get(), nnzL, Dv.get(), Dinv.get(), nk, fail.get(), active);
        };
        // C[rows >= c_lo, cols in [c_lo, ...)] -= L[:, k_lo .. k_lo + K) D L^T restricted to tile columns [tj0, tj1) / the first ncol columns
        auto update = [&](int k_lo, int K, int c_lo, cudaStream_t st, int tj0, int tj1, int ncol) {
            const int R = w.f - c_lo;
            if (R > 0) dense_syrk_sub_scaled(F + w.shift + c_lo + (size_t)k_lo * w.ld, front_stride, w.ld, Dv.get() + w.j0 + k_lo, nk,
                                             F + w.shift + c_lo + (size_t)c_lo * w.ld, front_stride, w.ld, R, K, batch, active, st, tj0, tj1, ncol);
        };
        const int nb0 = ((w.ws - 1) % MW_NB) + 1;        // first panel takes the remainder: every later boundary has the parity of ws (16-byte aligned row pairs)
        const int npan = 1 + (w.ws - nb0) / MW_NB;
        auto pstart = [&](int j) { return j == 0 ? 0 : nb0 + (j - 1) * MW_NB; };
        auto pwidth = [&](int j) { return j == 0 ? nb0 : MW_NB; };
        // The launch-by-launch timeline of config 3 (profiles/r02d_timeline_sparse_c3.txt) shows the factorisation bound by THIS chain (panel 75 us +
        // window update 50 us, 158 times = 20 ms; the far updates are shorter and hide behind it).  Panel j + 1 only needs its own columns
        // updated, so the window update is split: those columns first on the chain's stream, the columns of the later panels of the group on a
        // second stream beside panel j + 1.  Two updates that subtract into the same columns stay ordered: rest(j) runs after rest(j - 1)
        // (same stream) and first(j) waits for rest(j - 1).
        auto chain = [&](int ja, int jb, cudaStream_t st) {      // panels [ja, jb) of one group with their window updates
            const int gend = pstart(jb - 1) + pwidth(jb - 1);
            bool rest_pending = false;
            for (int j = ja; j < jb; j++) {
                panel(pstart(j), pwidth(j), st);
                const int r0 = pstart(j) + pwidth(j);
                if (j + 1 >= jb) continue;
                if (!aux2_stream) { update(pstart(j), pwidth(j), r0, st, 0, ceil_div(gend - r0, 128), gend - r0); continue; }      // 128 = tile width of the DMMA kernel
                const int w1 = pwidth(j + 1), c1 = r0 + w1;
                B200_CUDA(cudaEventRecord(ev_p2, st));
                B200_CUDA(cudaStreamWaitEvent(aux2_stream, ev_p2, 0));
                if (rest_pending) B200_CUDA(cudaStreamWaitEvent(st, ev_r2, 0));
                update(pstart(j), pwidth(j), r0, st, 0, 1, w1);                                                  // first(j): the columns of panel j + 1
                rest_pending = false;
                if (gend > c1) {
                    update(pstart(j), pwidth(j), c1, aux2_stream, 0, ceil_div(gend - c1, 128), gend - c1);       // rest(j): the later panels of the group
                    B200_CUDA(cudaEventRecord(ev_r2, aux2_stream));
                    rest_pending = true;
                }
            }
            if (rest_pending) B200_CUDA(cudaStreamWaitEvent(st, ev_r2, 0));
        };
        const int G = wide_group;
        chain(0, std::min(G, npan), stream);
        for (int ja = 0; ja < npan; ja += G) {
            const int jb = std::min(ja + G, npan), g0 = pstart(ja), g1 = pstart(jb - 1) + pwidth(jb - 1);
            const bool more = jb < npan;
            if (more && aux_stream) {
                const int jn = std::min(jb + G, npan), next_cols = pstart(jn - 1) + pwidth(jn - 1) - g1, tsplit = ceil_div(next_cols, 128);
                update(g0, g1 - g0, g1, stream, 0, tsplit, 0);
                B200_CUDA(cudaEventRecord(ev_col, stream));
                B200_CUDA(cudaStreamWaitEvent(aux_stream, ev_col, 0));
                chain(jb, jn, aux_stream);
                B200_CUDA(cudaEventRecord(ev_panel, aux_stream));
                update(g0, g1 - g0, g1, stream, tsplit, -1, 0);
                B200_CUDA(cudaStreamWaitEvent(stream, ev_panel, 0));
            } else {
                update(g0, g1 - g0, g1, stream, 0, -1, 0);
                if (more) chain(jb, std::min(jb + G, npan), stream);
            }
        }
        if (w.us > 0) { dim3 g(ceil_div(w.us, 128), w.us, batch);
            B200_LAUNCH(mfw_schur_kernel, g, 128, 0, stream, F, front_stride, w.ld, w.shift, w.ws, w.us, upd.get(), upd_total_w, w.off, active); }
    }
    for (const WStep& st : ws_steps) {           // inverses of the diagonal blocks of the supernodes the solves treat blocked
        if (st.kind != 1) continue;
        const int s2 = st.a, j0 = S.sup_ptr[s2], ws = S.sup_ptr[s2 + 1] - j0, us = S.rel_ptr[s2 + 1] - S.rel_ptr[s2];
        dim3 g(ceil_div(ws, wide_sb), batch);
        B200_LAUNCH(mfw_block_inverse_kernel, g, MW_T, sizeof(double) * (size_t)wide_sb * (wide_sb + 1), stream, ws, ws + us, S.Lp[j0], wide_sb, Lx.get(), nnzL,
                    Tcm.get(), Trm.get(), tinv_stride, (long long)st.b * wide_sb * wide_sb, active);
    }
}

void SparseLdltBatchedKKT::solve_wide(const double* rx, const double* ry, const double* rz, double* lx, double* ly, double* lz, const int* active) {
    const int nk = S.nk, sb = wide_sb;
    const size_t nnzL = std::max<size_t>(S.Li.size(), 1);
    dim3 gk(ceil_div(nk, 256), batch);
    B200_LAUNCH(ldlt_gather_rhs_kernel, gk, 256, 0, stream, d_perm.get(), n, S.pk, S.mk, rx, ry, rz, work.get(), active);
    const size_t fsm = sizeof(double) * (size_t)(MW_TS + sb), bsm = sizeof(double) * (size_t)(MW_TS + sb + 32);
    const int sr = std::min(64, sb);
    for (const WStep& st : ws_steps) {           // L y = b
        if (st.kind == 0) {
            dim3 g(ceil_div(st.b, MW_T / 32), batch);
            B200_LAUNCH(mfw_fwd_small_kernel, g, MW_T, 0, stream, d_wlist_s.get() + st.a, st.b, d_hdr.get(), d_Rp.get(), d_Rcol.get(), d_Rpos.get(), Lx.get(), nnzL, work.get(), nk, active);
            continue;
        }
        const int s2 = st.a, j0 = S.sup_ptr[s2], ws = S.sup_ptr[s2 + 1] - j0, us = S.rel_ptr[s2 + 1] - S.rel_ptr[s2], f = ws + us, lp0 = S.Lp[j0];
        { dim3 g(ceil_div(ws, MW_T / 32), batch);
          B200_LAUNCH(mfw_fwd_pull_kernel, g, MW_T, 0, stream, j0, ws, d_Rp.get(), d_Rcol.get(), d_Rpos.get(), Lx.get(), nnzL, work.get(), nk, active); }
        const int nblk = ceil_div(ws, sb);
        const long long toff = (long long)st.b * sb * sb;
        { dim3 g(1, batch); B200_LAUNCH(mfw_fwd_block_kernel, g, MW_TS, fsm, stream, j0, ws, f, lp0, 0, 0, sb, Lx.get(), nnzL, Tcm.get(), tinv_stride, toff, work.get(), nk, active); }
        for (int t = 0; t + 1 < nblk; t++) {
            const int after = ws - (t + 2) * sb;               // rows of the triangle behind the next block
            dim3 g(1 + (after > 0 ? ceil_div(after, sr) : 0), batch);
            B200_LAUNCH(mfw_fwd_block_kernel, g, MW_TS, fsm, stream, j0, ws, f, lp0, t * sb, sb, sb, Lx.get(), nnzL, Tcm.get(), tinv_stride, toff, work.get(), nk, active); }
    }
    B200_LAUNCH(ldlt_dscale_kernel, gk, 256, 0, stream, nk, Dinv.get(), work.get(), active);
    for (size_t i = ws_steps.size(); i-- > 0;) {   // L^T x = y
        const WStep& st = ws_steps[i];
        if (st.kind == 0) {
            dim3 g(ceil_div(st.b, MW_T / 32), batch);
            B200_LAUNCH(mfw_bwd_small_kernel, g, MW_T, 0, stream, d_wlist_s.get() + st.a, st.b, d_hdr.get(), d_Lp.get(), d_Li.get(), Lx.get(), nnzL, work.get(), nk, active);
            continue;
        }
        const int s2 = st.a, j0 = S.sup_ptr[s2], j1 = S.sup_ptr[s2 + 1] - 1, ws = j1 - j0 + 1, us = S.rel_ptr[s2 + 1] - S.rel_ptr[s2], f = ws + us, lp0 = S.Lp[j0];
        const int nblk = ceil_div(ws, sb);
        for (int t = nblk - 1; t >= 0; t--) {
            const int c0 = t * sb, cn = std::min(sb, ws - c0);
            dim3 g(cn, batch);
            B200_LAUNCH(mfw_bwd_block_kernel, g, MW_TS, bsm, stream, j0, ws, f, lp0, d_Li.get() + S.Lp[j1], c0, cn, sb, Lx.get(), nnzL, Trm.get(), tinv_stride,
                        (long long)st.b * sb * sb, work.get(), nk, wtmp.get(), wcounter.get(), active);
        }
    }
    B200_LAUNCH(ldlt_scatter_lhs_kernel, gk, 256, 0, stream, d_perm.get(), n, S.pk, S.mk, work.get(), lx, ly, lz, active);
}

SparseLdltBatchedKKT::SparseLdltBatchedKKT(SparseData* data, const int* user_perm, cudaStream_t st, int mode) : D(data) {
    batch = D->batch; n = D->n; p = D->p; m = D->m; stream = st;
    if (mode < 0 || mode > 3) throw std::runtime_error("sparse_ldlt: KKTMode must be 0..3");
    if (const char* e = getenv("B200_LDLT_LEVELS")) if (e[0] == '1') frontal = false;
    S.want_level_maps = !frontal;          // PK_to_L is read by the level-scheduled kernels only
    if (!S.analyse(D->P, D->AT, D->GT, user_perm, mode)) throw std::runtime_error(S.error);
    // value order of PKx: CSC order of the permuted matrix for the level kernels, ASSEMBLY order (grouped by front) for the
    // multifrontal kernels
    std::vector<int> order(S.PKi_rows.size());
    for (size_t q = 0; q < order.size(); q++) order[q] = (int)q;
    if (frontal) for (size_t t = 0; t < S.asm_q.size(); t++) order[S.asm_q[t]] = (int)t;
    auto compose = [&](const std::vector<int>& toK) { std::vector<int> r(toK.size()); for (size_t q = 0; q < toK.size(); q++) r[q] = toK[q] >= 0 ? order[S.K_to_PK[toK[q]]] : -1; return r; };
    upload(d_P_to_PK, compose(S.P_to_K)); upload(d_AT_to_PK, compose(S.AT_to_K)); upload(d_GT_to_PK, compose(S.GT_to_K));
    { std::vector<int> dg(S.diagPK.size()); for (size_t v = 0; v < dg.size(); v++) dg[v] = order[S.diagPK[v]]; upload(d_diagPK, dg); }
    upload(d_PK_to_L, S.PK_to_L); upload(d_PKp, S.PKp);
    upload(d_Lp, S.Lp); upload(d_Li, S.Li);
    if (!frontal) { S.build_row_view(nullptr); upload(d_Rp, S.Rp); upload(d_Rcol, S.Rcol); upload(d_Rpos, S.Rpos); }
    upload(d_level_cols, S.level_cols); upload(d_perm, S.perm);
    const size_t B = batch, nnzPK = S.PKi_rows.size(), nnzL = std::max<size_t>(S.Li.size(), 1);
    PKx.alloc(B * nnzPK); PKx.zero(st);
    P_diag.alloc(std::max<size_t>(B * n, 1));
    if (S.mode) {
        const int nxx = S.Kp[n];
        std::vector<int> tgt(nxx);
        for (int e = 0; e < nxx; e++) tgt[e] = order[S.K_to_PK[e]];
        upload(d_xx_P, S.xx_P); upload(d_xx_var, S.xx_var); upload(d_xx_ata, S.xx_ata); upload(d_xx_gtg, S.xx_gtg); upload(d_xx_target, tgt);
        upload(d_ata_ptr, S.ata.ptr); upload(d_ata_pa, S.ata.pa); upload(d_ata_pb, S.ata.pb);
        upload(d_gtg_ptr, S.gtg.ptr); upload(d_gtg_pa, S.gtg.pa); upload(d_gtg_pb, S.gtg.pb);
        AtA.alloc(std::max<size_t>(B * (size_t)S.ata.nnz(), 1));
        zinv.alloc(std::max<size_t>(B * m, 1)); dlt.alloc(B); crx.alloc(B * (size_t)n);
        zinv.zero(st); dlt.zero(st);
    }
    Lx.alloc(B * nnzL); Dv.alloc(B * S.nk); Dinv.alloc(B * S.nk); work.alloc(B * S.nk); fail.alloc(B); fail.zero(st);
    if (getenv("B200_MF_PROF")) { d_prof.alloc(8); d_prof.zero(st); }
    // multifrontal schedule (sparse_frontal.cuh)
    if (frontal) {
        std::vector<int> hdr((size_t)8 * S.nsup), crec((size_t)4 * std::max<size_t>(S.child_idx.size(), 1), 0);
        for (int s2 = 0; s2 < S.nsup; s2++) {
            const int j0 = S.sup_ptr[s2], j1 = S.sup_ptr[s2 + 1] - 1;
            int* h = &hdr[(size_t)8 * s2];
            h[0] = j0; h[1] = j1 - j0 + 1; h[2] = S.Lp[j1 + 1] - S.Lp[j1]; h[3] = S.Lp[j0];
            h[4] = S.asm_ptr[s2]; h[5] = S.asm_ptr[s2 + 1] - S.asm_ptr[s2]; h[6] = S.child_ptr[s2]; h[7] = S.child_ptr[s2 + 1] - S.child_ptr[s2];
        }
        for (size_t c = 0; c < S.child_idx.size(); c++) {
            const int cs = S.child_idx[c];
            crec[4 * c] = S.rel_ptr[cs + 1] - S.rel_ptr[cs]; crec[4 * c + 1] = S.rel_ptr[cs];
            crec[4 * c + 2] = (int)(unsigned)(S.upd_off[cs] & 0xffffffffll); crec[4 * c + 3] = (int)(S.upd_off[cs] >> 32);
        }
        upload(d_hdr, hdr); upload(d_crec, crec); upload(d_rel_idx, S.rel_idx); upload(d_asm_pos, S.asm_pos);
        d_upd_off.alloc(std::max<size_t>(S.upd_off.size(), 1));
        if (!S.upd_off.empty()) B200_CUDA(cudaMemcpy(d_upd_off.get(), S.upd_off.data(), S.upd_off.size() * sizeof(long long), cudaMemcpyHostToDevice));
        // few large QPs: spread each factorisation / solve over the whole GPU (sparse_wide.cuh)
        wide = (batch <= 4 && S.fmax >= 512) || (batch <= 16 && S.fmax >= 1024);      // few large QPs: spread every front over the GPU (a 1024-row front is 0.36 GFLOP: one CTA would need ~1 ms for it)
        if (wide) {      // every supernode keeps its own update slot in the wide schedule: stay with the stack discipline if that does not fit
            double slots = 0;
            for (int s2 = 0; s2 < S.nsup; s2++) { const double us = S.rel_ptr[s2 + 1] - S.rel_ptr[s2]; slots += us * us; }
            size_t free_b = 0, total_b = 0;
            B200_CUDA(cudaMemGetInfo(&free_b, &total_b));
            if (slots * sizeof(double) * B > 0.5 * (double)free_b) wide = false;
        }
        if (const char* e = getenv("B200_LDLT_WIDE")) wide = atoi(e) != 0;
        if (!wide) upd.alloc(B * (size_t)std::max<long long>(S.upd_total, 1));
        const size_t fpad = (size_t)((S.fmax + 1) & ~1);
        const size_t smem_cap = 200 * 1024, lcol_bytes = sizeof(double) * ((2 * fpad + 3) & ~(size_t)3);      // lcol (doubles) + relbuf (2 x fmax ints: mf_factor_kernel double-buffers it)
        const size_t big_scratch = sizeof(double) * (size_t)(2 * MF_TS * MF_NB + MF_NB * (MF_NB + 2));
        if (!wide && lcol_bytes + big_scratch > smem_cap) throw std::runtime_error("sparse_ldlt: a front of this size is not supported by this build");
        int fs = S.fmax;
        if (wide) { while (sizeof(double) * ((size_t)fs * fs + (size_t)(((fs + 1) & ~1) * 3 / 2 + 3)) > smem_cap) fs--; }      // per-level scratch: sized by the level's own fronts
        else while ((size_t)fs * fs * sizeof(double) + lcol_bytes > smem_cap) fs--;
        if (const char* e = getenv("B200_FRONT_SMEM_ROWS")) { const int v = atoi(e); if (v > 0) fs = std::min(fs, v); }     // tests: force the blocked HBM-front path
        front_smem_rows = fs;
        factor_smem = lcol_bytes + (size_t)fs * fs * sizeof(double);
        if (S.fmax > fs && !wide) {
            factor_smem = std::max(factor_smem, lcol_bytes + big_scratch);
            bigfront.alloc(B * (size_t)S.fmax * S.fmax); panel.alloc(B * 2 * (size_t)S.fmax * MF_NB);
        }
        if (wide) build_wide();
        solve_x_in_smem = (size_t)S.nk * sizeof(double) <= smem_cap;
        solve_smem = solve_x_in_smem ? (size_t)S.nk * sizeof(double) : 0;
        if (!wide) {
            allow_dynamic_smem(mf_factor_kernel, (size_t)((int)std::max<size_t>(factor_smem, 48 * 1024)));
            allow_dynamic_smem(mf_solve_kernel, (size_t)((int)std::max<size_t>(solve_smem, 48 * 1024)));
        }
        // streamed solve (mf_solve_ring_kernel): x + a double-buffered panel + row indices in shared memory
        ring_solve = false;
        if (!wide && solve_x_in_smem && !getenv("B200_LDLT_SIMPLE_SOLVE")) {
            const size_t xbytes = sizeof(double) * (size_t)((S.nk + 1) & ~1);
            const size_t rb = (size_t)((S.fmax + 3) & ~3);
            const size_t avail = smem_cap > xbytes + 2 * rb * sizeof(int) + 1024 ? smem_cap - xbytes - 2 * rb * sizeof(int) - 1024 : 0;
            size_t pb = std::min<size_t>(avail / (2 * sizeof(double)), 2048);      // 2 x 16 KB: enough to hide the L2 latency, small enough to keep the default smem carve-out
            if (const char* e = getenv("B200_RING_PB")) pb = std::min<size_t>(avail / (2 * sizeof(double)), (size_t)std::max(64, atoi(e)));
            pb = std::max<size_t>(pb, (size_t)S.fmax + 2);
            pb &= ~(size_t)1;
            if (pb >= (size_t)S.fmax + 2 && pb >= 256) {
                std::vector<int> sh;
                for (int s2 = 0; s2 < S.nsup; s2++) {
                    const int j0 = S.sup_ptr[s2], j1 = S.sup_ptr[s2 + 1] - 1, us = S.Lp[j1 + 1] - S.Lp[j1];
                    int c0 = 0; const int ws = j1 - j0 + 1;
                    while (c0 < ws) {
                        int c = 1;                             // widest block [c0, c0 + c) whose panel fits
                        auto panel = [&](int cc) { const long long nbelow = (ws - c0 - cc) + us; return (long long)cc * (cc - 1) / 2 + (long long)cc * nbelow; };
                        while (c0 + c < ws && panel(c + 1) <= (long long)pb) c++;
                        const int jb = j0 + c0, jl = jb + c - 1;
                        const int h[8] = {jb, c, S.Lp[jl + 1] - S.Lp[jl], S.Lp[jb], S.Lp[jl + 1] - S.Lp[jb], S.Lp[jl], 0, 0};
                        sh.insert(sh.end(), h, h + 8);
                        c0 += c;
                    }
                }
                ring_nblk = (int)sh.size() / 8; ring_pb = (int)pb; ring_rb = (int)rb;
                upload(d_shdr, sh);
                ring_smem = xbytes + 2 * pb * sizeof(double) + 2 * rb * sizeof(int);
                allow_dynamic_smem(mf_solve_ring_kernel, (size_t)((int)std::max<size_t>(ring_smem, 48 * 1024)));
                ring_solve = true;
            }
        }
    }
    scatter_static(7);
}
void SparseLdltBatchedKKT::copy_from(const SparseLdltBatchedKKT& o) {
    auto cp = [&](DevBuf<double>& d, const DevBuf<double>& s) { if (s.n) B200_CUDA(cudaMemcpyAsync(d.get(), s.get(), s.n * sizeof(double), cudaMemcpyDeviceToDevice, stream)); };
    cp(PKx, o.PKx); cp(P_diag, o.P_diag); cp(Lx, o.Lx); cp(Dv, o.Dv); cp(Dinv, o.Dinv);
    cp(AtA, o.AtA); cp(zinv, o.zinv); cp(dlt, o.dlt);
    if (wide && o.wide && Tcm.n == o.Tcm.n) { cp(Tcm, o.Tcm); cp(Trm, o.Trm); }
}
void SparseLdltBatchedKKT::scatter_static(int options) {   // kkt_full.hpp:212-251; update_data_impl of the condensed modes
    const int nnzPK = (int)S.PKi_rows.size();
    const bool elim_eq = S.mode & 1, elim_ineq = S.mode & 2;
    if ((options & 1) && S.mode == 0) {          // condensed modes re-add P at every factor (update_kkt_cost_scalings)
        if (D->P.nnz) { dim3 g(ceil_div(D->P.nnz, 256), batch);
            B200_LAUNCH(ldlt_scatter_kernel, g, 256, 0, stream, d_P_to_PK.get(), D->P.nnz, nnzPK, D->Px.get(), PKx.get()); }
        sparse_extract_diag(*D, P_diag.get(), stream);   // zeros where P has no stored diagonal
    }
    if ((options & 2) && D->AT.nnz) {
        if (elim_eq) update_AtA();
        else { dim3 g(ceil_div(D->AT.nnz, 256), batch);
            B200_LAUNCH(ldlt_scatter_kernel, g, 256, 0, stream, d_AT_to_PK.get(), D->AT.nnz, nnzPK, D->ATx.get(), PKx.get()); }
    }
    if ((options & 4) && D->GT.nnz && !elim_ineq) { dim3 g(ceil_div(D->GT.nnz, 256), batch);
        B200_LAUNCH(ldlt_scatter_kernel, g, 256, 0, stream, d_GT_to_PK.get(), D->GT.nnz, nnzPK, D->GTx.get(), PKx.get()); }
}
void SparseLdltBatchedKKT::update_AtA() {       // update_AT_A (kkt_eq_eliminated.hpp:223-245)
    const int nent = S.ata.nnz();
    if (nent == 0) return;
    dim3 g(ceil_div(nent, 128), batch);
    B200_LAUNCH(ldlt_gram_values_kernel, g, 128, 0, stream, d_ata_ptr.get(), d_ata_pa.get(), d_ata_pb.get(), nent, D->AT.nnz, D->ATx.get(), AtA.get());
}
void SparseLdltBatchedKKT::update_data(int options) { scatter_static(options); }

void SparseLdltBatchedKKT::factor(const double* delta, const double* x_reg, const double* z_reg, const int* active, int* ok) {   // sparse/kkt.hpp:83-105
    B200_ZONE("piqp::KKT::update_scalings_and_factor");
    const int nk = S.nk, nnzPK = (int)S.PKi_rows.size();
    const size_t nnzL = std::max<size_t>(S.Li.size(), 1);
    tic(T_ASSEMBLE);
    { dim3 g(ceil_div(nk, 256), batch);
      B200_LAUNCH(ldlt_set_diag_kernel, g, 256, 0, stream, d_diagPK.get(), n, S.pk, S.mk, S.mode ? 1 : 0, nnzPK, P_diag.get(), x_reg, delta, z_reg, PKx.get(), active); }
    if (S.mode) {
        const int nxx = S.Kp[n];
        { dim3 g(ceil_div(nxx, 128), batch);
          B200_LAUNCH(ldlt_cond_assemble_kernel, g, 128, 0, stream, nxx, d_xx_P.get(), d_xx_var.get(), d_xx_ata.get(), d_xx_gtg.get(), d_xx_target.get(),
                      d_gtg_ptr.get(), d_gtg_pa.get(), d_gtg_pb.get(), D->GT.d_colof.get(), n, m, D->P.nnz, D->GT.nnz, S.ata.nnz(), nnzPK,
                      D->Px.get(), D->GTx.get(), AtA.get(), x_reg, delta, z_reg, PKx.get(), active); }
        { dim3 g(ceil_div(std::max(m, 1), 256), batch);
          B200_LAUNCH(ldlt_store_scalings_kernel, g, 256, 0, stream, m, delta, z_reg, dlt.get(), zinv.get(), active); }
    }
    if (frontal) {
        toc(T_ASSEMBLE);
        tic(T_FACTOR);
        if (wide) factor_wide(active);
        else B200_LAUNCH(mf_factor_kernel, batch, MF_T, factor_smem, stream, make_mf(*this), PKx.get(), Lx.get(), Dv.get(), Dinv.get(), upd.get(), bigfront.get(), panel.get(), fail.get(), active);
        toc(T_FACTOR);
        B200_LAUNCH(ldlt_fail_to_ok_kernel, ceil_div(batch, 256), 256, 0, stream, fail.get(), active, ok, batch);
        return;
    }
    { dim3 g((unsigned)((nnzL + 255) / 256), batch); B200_LAUNCH(ldlt_zero_kernel, g, 256, 0, stream, Lx.get(), nnzL, active); }
    { dim3 g(ceil_div(nnzPK, 256), batch);
      B200_LAUNCH(ldlt_init_kernel, g, 256, 0, stream, d_PK_to_L.get(), nnzPK, nnzL, nk, PKx.get(), Lx.get(), Dv.get(), active); }
    toc(T_ASSEMBLE);
    tic(T_FACTOR);
    B200_LAUNCH(ldlt_clear_fail_kernel, ceil_div(batch, 256), 256, 0, stream, fail.get(), active, batch);
    const int nlev = (int)S.level_ptr.size() - 1;
    for (int l = 0; l < nlev; l++) {
        const int c0 = S.level_ptr[l], nc = S.level_ptr[l + 1] - c0;
        dim3 g(ceil_div(nc * 32, 128), batch);
        B200_LAUNCH(ldlt_level_kernel, g, 128, 0, stream, d_level_cols.get() + c0, nc, d_Lp.get(), d_Li.get(), d_Rp.get(), d_Rcol.get(), d_Rpos.get(),
                    nnzL, nk, Lx.get(), Dv.get(), Dinv.get(), fail.get(), active);
    }
    toc(T_FACTOR);
    B200_LAUNCH(ldlt_fail_to_ok_kernel, ceil_div(batch, 256), 256, 0, stream, fail.get(), active, ok, batch);
}

void SparseLdltBatchedKKT::solve(const double* rx, const double* ry, const double* rz, double* lx, double* ly, double* lz, const int* active) {   // sparse/kkt.hpp:107-176
    B200_ZONE("piqp::KKT::solve");
    const bool elim_eq = S.mode & 1, elim_ineq = S.mode & 2;
    tic(T_SOLVE);
    const double* srx = rx;
    if (S.mode) {      // :113-134: rhs_x += G^T (Z^-1 rhs_z) [ineq eliminated], += delta^-1 A^T rhs_y [eq eliminated]
        B200_CUDA(cudaMemcpyAsync(crx.get(), rx, sizeof(double) * (size_t)batch * n, cudaMemcpyDeviceToDevice, stream));
        if (elim_ineq && m > 0) spmv_rows(D->GT, D->GTx.get(), 1.0, rz, m, crx.get(), 1, zinv.get(), nullptr, 0, batch, active, stream);
        if (elim_eq && p > 0) spmv_rows(D->AT, D->ATx.get(), 1.0, ry, p, crx.get(), 1, nullptr, dlt.get(), 1, batch, active, stream);
        srx = crx.get();
    }
    solve_core(srx, ry, rz, lx, ly, lz, active);
    // :149-175: y = delta^-1 (A x - rhs_y), z = Z^-1 (G x - rhs_z)
    if (elim_eq && p > 0) spmv_cols(D->AT, D->ATx.get(), 1.0, lx, n, ly, ry, 1.0, nullptr, dlt.get(), 1, batch, active, stream);
    if (elim_ineq && m > 0) spmv_cols(D->GT, D->GTx.get(), 1.0, lx, n, lz, rz, 1.0, zinv.get(), nullptr, 0, batch, active, stream);
    toc(T_SOLVE);
}
// LDL^T solve of the (possibly condensed) system; blocks eliminated by the mode are never touched (S.pk / S.mk = 0)
void SparseLdltBatchedKKT::solve_core(const double* rx, const double* ry, const double* rz, double* lx, double* ly, double* lz, const int* active) {
    B200_ZONE("piqp::LDLt::solve_inplace");
    const int nk = S.nk;
    const size_t nnzL = std::max<size_t>(S.Li.size(), 1);
    if (frontal && wide) { solve_wide(rx, ry, rz, lx, ly, lz, active); return; }
    if (frontal) {
        if (ring_solve) B200_LAUNCH(mf_solve_ring_kernel, batch, MF_T, ring_smem, stream, make_mf(*this), d_shdr.get(), ring_nblk, ring_pb, ring_rb, Lx.get(), Dinv.get(),
                                    rx, ry, rz, lx, ly, lz, active);
        else B200_LAUNCH(mf_solve_kernel, batch, MF_T, solve_smem, stream, make_mf(*this), (int)solve_x_in_smem, Lx.get(), Dinv.get(), rx, ry, rz, lx, ly, lz, work.get(), active);
        return;
    }
    dim3 gk(ceil_div(nk, 256), batch);
    B200_LAUNCH(ldlt_gather_rhs_kernel, gk, 256, 0, stream, d_perm.get(), n, S.pk, S.mk, rx, ry, rz, work.get(), active);
    const int nlev = (int)S.level_ptr.size() - 1;
    for (int l = 1; l < nlev; l++) {     // level 0 rows have empty row patterns
        const int c0 = S.level_ptr[l], nc = S.level_ptr[l + 1] - c0;
        dim3 g(ceil_div(nc, 128), batch);
        B200_LAUNCH(ldlt_fwd_level_kernel, g, 128, 0, stream, d_level_cols.get() + c0, nc, d_Rp.get(), d_Rcol.get(), d_Rpos.get(), nnzL, nk, Lx.get(), work.get(), active);
    }
    B200_LAUNCH(ldlt_dscale_kernel, gk, 256, 0, stream, nk, Dinv.get(), work.get(), active);
    for (int l = nlev - 2; l >= 0; l--) {   // the top level has empty columns
        const int c0 = S.level_ptr[l], nc = S.level_ptr[l + 1] - c0;
        dim3 g(ceil_div(nc, 128), batch);
        B200_LAUNCH(ldlt_bwd_level_kernel, g, 128, 0, stream, d_level_cols.get() + c0, nc, d_Lp.get(), d_Li.get(), nnzL, nk, Lx.get(), work.get(), active);
    }
    B200_LAUNCH(ldlt_scatter_lhs_kernel, gk, 256, 0, stream, d_perm.get(), n, S.pk, S.mk, work.get(), lx, ly, lz, active);
}
void SparseLdltBatchedKKT::eval_P_x(double alpha, const double* x, double* z, const int* active) { spmv_sym_upper(D->P, D->Px.get(), alpha, x, z, batch, active, stream); }
void SparseLdltBatchedKKT::eval_A(double an, double at, const double* xn, const double* xt, double* zn, double* zt, const int* active) {
    if (p > 0) spmv_cols(D->AT, D->ATx.get(), an, xn, n, zn, nullptr, 0.0, nullptr, nullptr, 0, batch, active, stream);
    spmv_rows(D->AT, D->ATx.get(), at, xt, p, zt, 0, nullptr, nullptr, 0, batch, active, stream);
}
void SparseLdltBatchedKKT::eval_G(double an, double at, const double* xn, const double* xt, double* zn, double* zt, const int* active) {
    if (m > 0) spmv_cols(D->GT, D->GTx.get(), an, xn, n, zn, nullptr, 0.0, nullptr, nullptr, 0, batch, active, stream);
    spmv_rows(D->GT, D->GTx.get(), at, xt, m, zt, 0, nullptr, nullptr, 0, batch, active, stream);
}
void SparseLdltBatchedKKT::extract_P_diag(double* out) { sparse_extract_diag(*D, out, stream); }
void SparseLdltBatchedKKT::print_info() const {
    printf("b200 sparse_ldlt backend: n_kkt = %d, nnz(KKT upper) = %zu, nnz(L) = %.0f, etree levels = %zu, supernodes = %d, largest front = %d, factor flops = %.3g\n",
           S.nk, S.PKi_rows.size(), S.nnzL(), S.level_ptr.size() - 1, S.nsup, S.fmax, S.factor_flops());
}

}  // namespace b200
