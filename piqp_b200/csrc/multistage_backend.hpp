// piqp_b200/csrc/multistage_backend.hpp -- batched block-tridiagonal-arrow KKT backend (sparse_multistage).
//
// Replaces sparse::MultistageKKT (include/piqp/sparse/multistage_kkt.hpp:41-1816) for a batch of QPs that share one
// sparsity pattern:
//   host, once     : MsStructure::detect  == extract_arrow_structure (:420-597)  (integer logic, restated)
//                    slot maps / contribution lists == utri_to_kkt + transpose_to_block_mat + block_syrk_ln bookkeeping
//   device, per it.: ms_assemble_kernel    == block_gemm_nd + block_syrk_ln_calc + populate_kkt_fac (:180-215, 833-1219)
//                    ms_factor_kernel      == factor_kkt (:1253-1352)          one CTA per instance walking the chain in smem
//                    ms_solve_kernel       == solve_llt_in_place (:1709-1816)  one CTA per instance, inverse-block mat-vecs
//                    SpMV (sparse_data.cu) == block_t_gemv_* / block_symv_l / BlockVec::assign/load (:291-383,1355-1706)
#pragma once
#include <string>
#include "kkt_backend.hpp"
#include "sparse_data.hpp"
#include <vector>

namespace b200 {

struct MsBlock { int start, diag, off; };

struct MsStructure {   // host-side description shared by all instances
    int n = 0, N = 0, w = 0, dmax = 0, omax = 0, total = 0, total_inv = 0;
    std::vector<MsBlock> bi;
    std::vector<int> offD, offB, offE, offI;       // offsets of D_i (d x d), B_i (o x d), E_i (w x d) in the block storage; of inv(L_i) in the inverse storage
    std::vector<int> blk_of;
    std::vector<int> P_slot, diag_slot;            // P nnz -> slot ; variable -> slot of its diagonal element
    std::vector<int> a_ptr, a_qa, a_qb;            // AtA contributions grouped by slot
    std::vector<int> g_ptr, g_qa, g_qb, g_row;     // GtG contributions grouped by slot (+ row of G for the weight)
    std::string error;
    bool detect(const Pattern& P, const Pattern& AT, const Pattern& GT);
    int slot(int i, int j) const;                  // lower element (i >= j) -> slot or -1
    double factor_flops() const;                   // the reference's own cost model (:397-418)
    double factor_bytes() const;
    double solve_flops() const;
    double solve_bytes() const;
};

class MultistageBatchedKKT : public BatchedKKT {
public:
    MultistageBatchedKKT(SparseData* data, cudaStream_t st);   // throws std::runtime_error if the pattern does not fit the block structure
    void update_data(int options) override;
    void factor(const double* delta, const double* x_reg, const double* z_reg, const int* active, int* ok) override;
    void solve(const double* rx, const double* ry, const double* rz, double* lx, double* ly, double* lz, const int* active) override;
    void eval_P_x(double alpha, const double* x, double* z, const int* active) override;
    void eval_A(double an, double at, const double* xn, const double* xt, double* zn, double* zt, const int* active) override;
    void eval_G(double an, double at, const double* xn, const double* xt, double* zn, double* zt, const int* active) override;
    void extract_P_diag(double* P_diag) override;
    void print_info() const override;
    bool graph_capturable() const override;
    bool factor_never_fails() const override { return true; }      // multistage_kkt.hpp:218: update_scalings_and_factor always returns true
    double factor_flops() const override { return S.factor_flops(); }
    double factor_bytes() const override { return S.factor_bytes(); }
    double solve_flops() const override { return S.solve_flops(); }
    double solve_bytes() const override { return S.solve_bytes(); }
    void copy_from(const MultistageBatchedKKT& o);

    SparseData* D;
    MsStructure S;
    DevBuf<int> d_meta;                 // start[N] diag[N] off[N] offD[N] offB[N] offE[N] offI[N]
    DevBuf<int> d_P_slot, d_diag_var;   // P nnz -> slot ; slot -> variable index of a diagonal element or -1
    DevBuf<int> d_a_ptr, d_a_qa, d_a_qb, d_g_ptr, d_g_qa, d_g_qb, d_g_row;
    DevBuf<double> Pblk, AtAblk, fac, Linv;   // [batch][total] x3, [batch][total_inv]
    DevBuf<double> zinv, delta, work_z;
    size_t factor_smem = 0, solve_smem = 0, chain_solve_smem = 0;
    bool warp_chain = false;            // fronts of <= 32 rows: one warp walks the chain (multistage_chain.cuh)
    int chain_slot = 0, chain_rp = 16;  // doubles per ring slot of msw_solve_kernel; ceil(max front rows / 2)
    DevBuf<double> packets;             // [batch][pk_stride] solve packets (see multistage_chain.cuh)
    size_t pk_stride = 0;
    // ---- parallel-in-horizon partition (multistage_partition.cuh): K runs of stages separated by K-1 separator stages
    int part_K = 1;                     // 1 = off
    std::vector<int> part_bounds, part_sep, part_dsep;     // [2K] stage ranges of the runs; separator stages; per stage: class of the separator on the left of its run (0: none)
    int part_rn = 0, part_rtotal = 0, part_rslot = 0, part_rrp = 16, part_seg_len = 0, part_dsep_max = 8, part_spike_slot = 2, part_solve_slot = 2;
    size_t part_rpk_stride = 0, part_seg_smem = 0, part_spike_smem = 0, part_rsolve_smem = 0, part_fused_smem = 0;      // part_fused_smem > 0: one launch per solve (msp_solve_fused_kernel)
    int part_fused_run = 0, part_fused_slot = 0;
    DevBuf<int> d_rmeta, d_part;
    DevBuf<double> rfac, rpackets, carry, zbuf, xred;
private:
    void load_P();
    void compute_AtA();
    void plan_partition(const std::vector<int>& cls);
    void build_partition(const std::vector<int>& cls, cudaStream_t st);
    struct MsPart make_part() const;
    void factor_partitioned(const struct MsDev& dv, const int* active);
    void solve_partitioned(double* lx, const int* active);
};

}  // namespace b200
