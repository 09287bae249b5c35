// piqp_b200/csrc/sparse_ldlt_backend.hpp -- batched general sparse KKT backend (sparse_ldlt, KKTMode FULL).
//
// Replaces sparse::KKT<T,I,KKT_FULL> (include/piqp/sparse/kkt.hpp:31-250) with its helpers
//   KKTImpl<FULL>::create_kkt_matrix / update_kkt_* / update_data_impl   include/piqp/sparse/kkt_full.hpp:39-251
//   AMDOrdering (Eigen::AMDOrdering, third party)                        include/piqp/sparse/ordering.hpp:59-125
//   permute_sparse_symmetric_matrix                                      include/piqp/sparse/utils.hpp:31-128
//   LDLt::factorize_symbolic / factorize_numeric / solve_inplace         include/piqp/sparse/ldlt.hpp:42-218
// for a batch of QPs sharing one sparsity pattern.
//
// Host, once per pattern: KKT pattern -> fill-reducing ordering (own quotient-graph minimum degree; any valid
// permutation yields the same solve() results to rounding) -> permuted upper pattern -> elimination tree, pattern of L,
// etree LEVEL SETS.  Device, per iteration: the reference's up-looking row-by-row LDL^T is inherently serial, so the
// numeric factorisation is restated as a LEFT-LOOKING column algorithm scheduled by etree levels: all columns of one
// level are independent, one warp per (column, instance) gathers the updates of its descendants (binary search into
// the target column's sorted row list), then scales by D_j.  Triangular solves are level-scheduled gathers (row view for
// L, column view for L^T): deterministic, no atomics.  Zero pivots report failure like ldlt.hpp:161.
#pragma once
#include <string>
#include "kkt_backend.hpp"
#include "sparse_data.hpp"

namespace b200 {

struct LdltSymbolic {   // host
    int n = 0, p = 0, m = 0, nk = 0;
    int mode = 0, pk = 0, mk = 0;                 // KKTMode (kkt_fwd.hpp:15-21); sizes of the y / z blocks KEPT in the KKT (0 when eliminated)
    // condensed modes: contribution lists of upper(A^T A) / upper(G^T Z^-1 G) and, per entry of the top-left block of K
    // (K value index < Kp[n]), where its summands come from (kkt_all_eliminated.hpp:108-160)
    struct Gram { std::vector<int> colp, rows, ptr, pa, pb; int nnz() const { return (int)rows.size(); } };
    Gram ata, gtg;
    std::vector<int> xx_P, xx_var, xx_ata, xx_gtg; // P value index / variable (diagonal entries) / ata entry / gtg entry, -1 = none
    static void build_gram(const Pattern& MT, Gram& g);
    void build_row_view(const std::vector<char>* skip_sup);      // fills Rp / Rcol / Rpos
    std::vector<int> perm, iperm;                 // perm[new] = old ; iperm[old] = new   (ordering.P / P_inv)
    // unpermuted KKT (upper CSC) and the maps of kkt_full.hpp:39-170
    std::vector<int> Kp, Ki;
    std::vector<int> P_to_K, AT_to_K, GT_to_K;    // value index of P_utri / AT / GT -> value index of K
    std::vector<int> K_to_PK;                     // PKi: value index of K -> value index of the permuted upper matrix
    std::vector<int> PKp, PKi_rows;               // permuted upper CSC pattern
    std::vector<int> diagPK;                      // variable (unpermuted KKT index) -> value index of its diagonal in PK
    // L (unit lower, CSC, sorted rows) and its row view
    std::vector<int> Lp, Li, etree, level, level_ptr, level_cols;
    std::vector<int> Rp, Rcol, Rpos;              // row j: columns k < j with L(j,k) != 0 and the position of L(j,k) in column k
    std::vector<int> PK_to_L;                     // value index of PK -> position in L (off-diagonal) or -(j+1) for the diagonal of column j
    // ---- supernodal multifrontal schedule (columns are postordered, so a supernode is a run of consecutive columns whose
    //      L patterns are nested: struct(L_j) = {j+1..j1} u U, U = struct(L_j1)); front of supernode s = rows {j0..j1} u U
    int nsup = 0, fmax = 0;                       // number of supernodes, largest front
    std::vector<int> sup_ptr;                     // [nsup+1] first column of each supernode
    std::vector<int> child_ptr, child_idx;        // children (supernodes) of each supernode, in increasing order
    std::vector<int> rel_ptr, rel_idx;            // per supernode: position of each of ITS update rows in its PARENT's front
    std::vector<int> asm_ptr, asm_q, asm_pos;     // per supernode: entries of the permuted matrix (value index, offset row + col * f in the front)
    std::vector<long long> upd_off;               // per supernode: offset of its update matrix (us x us, lower, ld = us) on the per-instance stack
    long long upd_total = 0;                      // stack size in doubles
    std::string error;
    bool want_level_maps = true;                  // build PK_to_L (scatter map of the level-scheduled kernels)
    bool analyse(const Pattern& P, const Pattern& AT, const Pattern& GT, const int* user_perm, int mode = 0);
    double nnzL_exact = -1.0, flops_exact = -1.0; // of the unpadded pattern (what the reference's LDLt would store / compute)
    double nnzL() const { return nnzL_exact >= 0 ? nnzL_exact : (Lp.empty() ? 0.0 : (double)Lp.back()); }
    double factor_flops() const;                  // sum_j (c_j^2 + 2 c_j), SURVEY 8(d)
};

std::vector<int> minimum_degree_ordering(int n, const std::vector<int>& colptr, const std::vector<int>& rowidx);   // pattern of upper(A), A symmetric

class SparseLdltBatchedKKT : public BatchedKKT {
public:
    SparseLdltBatchedKKT(SparseData* data, const int* user_perm, cudaStream_t st, int mode = 0);
    ~SparseLdltBatchedKKT() override;
    void update_data(int options) override;
    void factor(const double* delta, const double* x_reg, const double* z_reg, const int* active, int* ok) override;
    void solve(const double* rx, const double* ry, const double* rz, double* lx, double* ly, double* lz, const int* active) override;
    void eval_P_x(double alpha, const double* x, double* z, const int* active) override;
    void eval_A(double an, double at, const double* xn, const double* xt, double* zn, double* zt, const int* active) override;
    void eval_G(double an, double at, const double* xn, const double* xt, double* zn, double* zt, const int* active) override;
    void extract_P_diag(double* P_diag) override;
    void print_info() const override;
    // the whole-GPU schedule forks onto auxiliary streams (look-ahead) and joins them again before every factor / solve returns, so it can be
    // captured like the dense Cholesky's side streams; its thousands of launches per iteration are what a graph replay is for (B200_WIDE_NO_GRAPH=1: stepwise)
    bool graph_capturable() const override { return !wide || wide_graph; }
    bool wide_graph = false;      // set by build_wide(): schedules with many HBM fronts (thousands of short dependent launches per iteration) replay as a graph;
                                  // one or two huge fronts (config 3) run 3 % faster stepwise (stream priorities of the look-ahead chain)
    double factor_flops() const override { return S.factor_flops(); }
    double factor_bytes() const override { return 12.0 * S.nnzL() + 12.0 * (double)S.PKi_rows.size(); }
    double solve_flops() const override { return 4.0 * S.nnzL() + S.nk; }
    double solve_bytes() const override { return 2.0 * 12.0 * S.nnzL() + 5.0 * 8.0 * S.nk; }
    void copy_from(const SparseLdltBatchedKKT& o);

    SparseData* D;
    LdltSymbolic S;
    DevBuf<int> d_P_to_PK, d_AT_to_PK, d_GT_to_PK, d_diagPK, d_PK_to_L, d_PKp;
    DevBuf<int> d_Lp, d_Li, d_Rp, d_Rcol, d_Rpos, d_level_cols, d_perm;
    DevBuf<double> PKx;        // [batch][nnz(PK)] permuted KKT values (upper)
    DevBuf<double> P_diag;     // [batch][n]
    DevBuf<double> Lx, Dv, Dinv;   // [batch][nnz(L)], [batch][nk] x2
    DevBuf<double> work;       // [batch][nk] permuted rhs / solution
    DevBuf<int> fail;
    DevBuf<long long> d_prof;      // B200_MF_PROF=1: phase clocks of mf_factor_kernel's CTA 0 (diagnostics)
    // condensed modes (sparse_ldlt_eq_cond / _ineq_cond / _cond)
    DevBuf<int> d_xx_P, d_xx_var, d_xx_ata, d_xx_gtg, d_xx_target, d_ata_ptr, d_ata_pa, d_ata_pb, d_gtg_ptr, d_gtg_pa, d_gtg_pb;
    DevBuf<double> AtA;        // [batch][nnz(upper(A^T A))], recomputed on update_data(A)
    DevBuf<double> zinv, dlt;  // [batch][m] z_reg^-1 and [batch] delta of the last factor() (m_z_reg_inv / m_delta, sparse/kkt.hpp:36-38)
    DevBuf<double> crx;        // [batch][n] condensed rhs_x
    // multifrontal path
    bool frontal = true;       // B200_LDLT_LEVELS=1 selects the level-scheduled simplicial kernels instead
    int front_smem_rows = 0;   // fronts up to this many rows live in shared memory, larger ones in `bigfront`
    size_t factor_smem = 0, solve_smem = 0;
    bool solve_x_in_smem = true;
    DevBuf<int> d_hdr, d_crec, d_rel_idx, d_asm_pos, d_shdr;
    bool ring_solve = false;   // mf_solve_ring_kernel: L panels streamed through a cp.async double buffer
    int ring_nblk = 0, ring_pb = 0, ring_rb = 0;
    size_t ring_smem = 0;
    DevBuf<long long> d_upd_off;
    DevBuf<double> upd, bigfront, panel;   // [batch][upd_total], [batch][fmax^2] and [batch][2 fmax NB] (only if fmax > front_smem_rows)
    // ---- whole-GPU ("wide") schedule for few large QPs (sparse_wide.cuh): supernodes by etree level, HBM fronts spread over the GPU
    bool wide = false;
    int wide_sb = 128;         // column block of the blocked supernodal solves (<= 128)
    int wide_group = 4;        // panels per group of the two-level blocked LDL^T (far updates contract over 64 * wide_group columns)
    struct WStep { int kind, a, b, c; };      // kind 0: narrow / shared-memory group (list offset a, count b, factor: fpad c); kind 1: wide supernode (factor: index into wfronts, solve: supernode)
    struct WFront { int s, j0, ws, us, f, ld, shift, lp0, ab, an, pull_begin, nchild; long long off; };
    std::vector<WStep> wf_steps, ws_steps;
    std::vector<WFront> wfronts;
    std::vector<size_t> wf_smem;              // dynamic shared memory of each factor step of kind 0
    DevBuf<int> d_wlist_f, d_wlist_s, d_crecw, d_pull_ptr, d_pull_child, d_pull_cc;
    DevBuf<long long> d_upd_off_w;
    DevBuf<double> wtmp, Tcm, Trm;   // dot products of one backward block; inverses of the diagonal blocks of the wide supernodes (column- / row-major)
    long long tinv_stride = 0;
    DevBuf<unsigned> wcounter;
    long long upd_total_w = 0, front_stride = 0;
    cudaStream_t aux_stream = nullptr;        // look-ahead: panel k+1 runs here while the rest of trailing update k runs on `stream`
    cudaEvent_t ev_col = nullptr, ev_panel = nullptr;
    cudaStream_t aux2_stream = nullptr;       // the part of a window update that the next panel does not need runs here, beside that panel
    cudaEvent_t ev_p2 = nullptr, ev_r2 = nullptr;
private:
    void build_wide();
    void factor_wide(const int* active);
    void solve_wide(const double* rx, const double* ry, const double* rz, double* lx, double* ly, double* lz, const int* active);
    void scatter_static(int options);
    void update_AtA();
    void solve_core(const double* rx, const double* ry, const double* rz, double* lx, double* ly, double* lz, const int* active);
};

}  // namespace b200
