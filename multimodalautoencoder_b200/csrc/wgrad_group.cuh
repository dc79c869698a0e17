// Grouped weight-gradient GEMM for the small-width configs: EVERY dW = in^T . delta of a train step in ONE
// persistent tcgen05 launch (multimodal_autoencoder.py:411 / :443 -- the gradient half of opt_step).
//
//   problem q:  C_q[M_q, N_q] = A_q^T . B_q   with A_q stored [batch, M_q] and B_q stored [batch, N_q]
//               (the contraction runs over the batch, so both operands are MN-major for tcgen05)
//   work item = (problem, 128 x 128 output tile, batch slice); the items of all problems form one static schedule over
//               the 148 persistent CTAs.  Each item writes its partial tile into a split-K slice; the slices are summed
//               in a fixed order by grad_assemble_kernel (kernels.cuh), so the result is deterministic.
//
// Same warp roles, smem ring and epilogue staging as gemm_tc_kernel<128, true, true> (gemm_tc_kernel.cuh); what is new
// is the problem table (tensor maps per problem in the kernel parameters) and the schedule across problems, which is
// what removes one launch + one reduction launch per layer from the step.
#pragma once
#include <vector>

#include "gemm_tc.cuh"

namespace mmae {

constexpr int WG_MAX_PROBLEMS = 12;     // 12 x (2 tensor maps + 48 B) stays below the 4 KB kernel-parameter space
constexpr int WG_BN = 128;

struct WgProblem {
  CUtensorMap tmA, tmB;
  int M, N;                 // output rows / columns
  int m_blocks, n_blocks;
  int tile0;                // first tile of this problem in the flattened (problem, m, n) list
  int a3d, b3d;             // operand loaded through the 3-D MN-major map (width % 32 == 0)
  float* ws;                // slices [splits][M][N]
};

struct WgParams {
  WgProblem pr[WG_MAX_PROBLEMS];
  int nprob, tiles;         // tiles = sum of m_blocks * n_blocks
  int splits;
  int64_t K, k_per_split;   // batch, rows per slice (multiple of TC_BK)
};

// one weight gradient of the step, as the engine describes it
struct WgDesc {
  const float* A; int64_t lda; int M;      // in  [batch, M]
  const float* B; int64_t ldb; int N;      // delta [batch, N]
};

struct WgPlan { int splits; int64_t k_per_split; int grid; int tiles; };

inline WgPlan wg_plan(const std::vector<WgDesc>& d, int64_t K, int num_sms) {
  WgPlan pl; pl.tiles = 0;
  for (const WgDesc& w : d) pl.tiles += ((w.M + TC_BM - 1) / TC_BM) * ((w.N + WG_BN - 1) / WG_BN);
  const int64_t kblocks = (K + TC_BK - 1) / TC_BK;
  // as many batch slices as fill the persistent grid once, at least 8 k-blocks each
  int splits = (int)std::max<int64_t>(1, std::min<int64_t>(num_sms / std::max(pl.tiles, 1), kblocks / 8));
  splits = std::max(1, std::min(splits, 64));
  const int64_t kb_per = (kblocks + splits - 1) / splits;
  pl.k_per_split = kb_per * TC_BK;
  pl.splits = (int)((kblocks + kb_per - 1) / kb_per);
  const int64_t total = (int64_t)pl.tiles * pl.splits;
  pl.grid = (int)std::min<int64_t>(total, num_sms);
  return pl;
}

inline bool wg_eligible(const WgDesc& w) {
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  return al(w.A) && al(w.B) && !(w.lda & 3) && !(w.ldb & 3) && !(w.M & 3) && !(w.N & 3) && w.M >= 32 && w.N >= 32;
}

// defined in wgrad_group.cu
cudaError_t wgrad_group_launch(const WgParams& p, int grid, cudaStream_t st);

}  // namespace mmae
