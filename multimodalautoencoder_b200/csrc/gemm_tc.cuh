// tcgen05 / TMEM / TMA GEMM for sm_100a, kind::tf32 (fp32 storage, fp32 accumulation in TMEM).
//
//   C[M,N] = opA(A) . opB(B) with the fused MMAE epilogue (common.cuh).
//   A_MN: A stored [K,M] (M contiguous -> MN-major operand, wgrad);  else [M,K] (K-major).
//   B_MN: B stored [K,N] (N contiguous -> MN-major operand, y = x.W); else [N,K] (K-major, dgrad / tied).
//
// Structure (one CTA per SM, persistent over a static tile schedule):
//   warp 0      TMA producer: cp.async.bulk.tensor.2d -> 128B-swizzled smem ring, mbarrier expect_tx
//   warp 1      MMA issuer: one lane issues tcgen05.mma.cta_group::1.kind::tf32 (M=128, N=BN, K=8),
//               tcgen05.commit frees the smem stage / publishes the accumulator
//   warps 2..9  epilogue: tcgen05.ld 32x32b.x32 (two warps per TMEM lane quadrant, alternating 32-column
//               chunks), transposed through a padded smem staging tile so that every global access of the
//               fused bias/activation/dropout/loss/act' epilogue (C, target, saved activations) is a
//               coalesced 128-byte row segment
//   TMEM: 2 accumulator stages x BN fp32 columns, so the epilogue of tile t overlaps the MMAs of t+1.
// Tails: TMA zero-fills out-of-bounds rows / columns / K; the epilogue bounds-checks its stores.
// Requirements (checked by tc_gemm_eligible): lda, ldb, ldc multiples of 4 floats, 16-byte aligned bases.
#pragma once
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "gemm_simt.cuh"   // GemmArgs

namespace mmae {

constexpr int TC_BM = 128;
constexpr int TC_BK = 32;                 // 32 fp32 = 128 bytes = one swizzle row
constexpr int TC_UMMA_K = 8;              // kind::tf32: 32 bytes of K per instruction
constexpr int TC_THREADS = 320;           // 10 warps: TMA, MMA, 8 epilogue
constexpr int TC_EPI_WARPS = 8;
constexpr int TC_STAGE_LD = 33;           // epilogue staging row stride (floats): conflict-free both ways
constexpr int TC_EPI_WARP0 = 2;

template <int BN> struct TcCfg {
  static constexpr int kStages = BN == 256 ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int kABytes = TC_BM * TC_BK * 4;          // 16 KB
  static constexpr int kBBytes = BN * TC_BK * 4;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = 2 * BN;                   // 128 / 256 / 512 (power of two)
  static constexpr int kEpiBytes = TC_EPI_WARPS * 32 * TC_STAGE_LD * 4;   // 33 792 B
  static constexpr int kSmemBytes = kStages * kStageBytes + kEpiBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

struct TcParams {
  CUtensorMap tmA, tmB;
  int64_t M, N, K;
  float* C; int64_t ldc;
  int m_blocks, n_blocks, splits;
  int64_t k_per_split;          // multiple of TC_BK
  int64_t split_stride;         // elements between split-K slices of C (0 when splits == 1)
  int a3d, b3d;                 // MN-major operand loaded through a 3-D map (one TMA per stage)
  Epilogue ep;
  NoiseView nz;                 // two-SM kernel only: block-mask + zero noise applied to the A tile in shared memory
  int nz_aligned;               // every modality boundary is a multiple of 32 columns (one modality per 32-column chunk)
  CUtensorMap tmC;              // two-SM kernel only: [32 x 32] boxes over C for the row-layout TMA-store epilogue
  int tma_epi;                  // 1: bias/act, dgrad and loss epilogues keep thread = row and leave through tmC
};

// ------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(f);
  }
  return fn;
}

// 2-D fp32 tensor map over a row-major [rows, cols] matrix (cols contiguous), 128B swizzle.
// mn_major operands (32-bit elements) must use the 128B swizzle with 32-byte atomicity
// ("for mn-major tf32 operands, SW128_32B is the only available smem layout").
inline bool make_tmap(CUtensorMap* tm, const float* base, int64_t rows, int64_t cols, int64_t ld,
                      int box_cols, int box_rows, bool mn_major = false) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  static int use_tf32_type = -1;     // MMAE_TMA_TF32=0 loads raw fp32 (the MMA then truncates to tf32)
  if (use_tf32_type < 0) { const char* ev = getenv("MMAE_TMA_TF32"); use_tf32_type = (ev && ev[0] == '0') ? 0 : 1; }
  CUresult r = enc(tm, use_tf32_type ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// 3-D view of an MN-major operand stored [K rows, MN cols]: (32 mn, K, MN/32 chunks) so that ONE TMA instruction
// lands all 32-column chunks of a stage ([chunk][k][32] in smem, 128B swizzle with 32-byte atoms).  Needs MN % 32 == 0.
inline bool make_tmap_mn3d(CUtensorMap* tm, const float* base, int64_t k_rows, int64_t mn, int64_t ld, int box_k,
                           int box_chunks) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return false;
  static int use_tf32_type = -1;
  if (use_tf32_type < 0) { const char* ev = getenv("MMAE_TMA_TF32"); use_tf32_type = (ev && ev[0] == '0') ? 0 : 1; }
  cuuint64_t dims[3] = {32, (cuuint64_t)k_rows, (cuuint64_t)(mn / 32)};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 4, 128};
  cuuint32_t box[3] = {32, (cuuint32_t)box_k, (cuuint32_t)box_chunks};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(tm, use_tf32_type ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                   const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// plain-fp32 2-D map with [32 x 32] boxes (epilogue loads / stores; TMA clips rows and columns out of range)
inline bool make_tmap_io(CUtensorMap* tm, const float* base, int64_t rows, int64_t cols, int64_t ld) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

inline bool tc_gemm_eligible(bool ta, bool tb, const GemmArgs& g, bool allow_noise = false) {
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (g.noise.enabled && !allow_noise) return false; // one-SM family: the noisy operand is materialised first
  if ((g.lda & 3) || (g.ldb & 3) || (g.ldc & 3)) return false;
  if (!al(g.A) || !al(g.B) || !al(g.C)) return false;
  if (g.M < 32 || g.N < 32 || g.K < 32 || (g.N & 3)) return false;   // smaller dims ride on TMA zero-fill; below 32 the CUDA-core family wins
  if (g.ep.target && (g.ep.ldt & 3)) return false;
  (void)ta; (void)tb;
  return true;
}

inline int tc_pick_bn(int64_t N) { return N > 128 ? 256 : (N > 64 ? 128 : 64); }

struct TcPlan { int bn; int m_blocks, n_blocks, splits; int64_t k_per_split; int grid; };

inline TcPlan tc_plan(const GemmArgs& g, int num_sms, int max_splits) {
  TcPlan pl;
  pl.bn = tc_pick_bn(g.N);
  pl.m_blocks = (int)((g.M + TC_BM - 1) / TC_BM);
  pl.n_blocks = (int)((g.N + pl.bn - 1) / pl.bn);
  int64_t tiles = (int64_t)pl.m_blocks * pl.n_blocks;
  int64_t kblocks = (g.K + TC_BK - 1) / TC_BK;
  // split-K (wgrad only): pick the split count that best fills whole waves of the persistent grid, keeping at
  // least 32 k-blocks per split (8 while the grid is still below one wave) and preferring fewer splits when the gain is below 3 % (slices cost HBM traffic)
  int splits = 1;
  if (max_splits > 1) {
    double best = 0.0;
    for (int s = 1; s <= max_splits; ++s) {
      if (s > 1 && kblocks / s < (tiles * s <= num_sms ? 8 : 32)) break;    // small grids: shorter slices beat idle SMs
      const int64_t total = tiles * s;
      const int64_t waves = (total + num_sms - 1) / num_sms;
      const double eff = (double)total / (double)(waves * num_sms);
      if (eff > best + 0.03) { best = eff; splits = s; }
    }
  }
  int64_t kb_per = (kblocks + splits - 1) / splits;
  pl.k_per_split = kb_per * TC_BK;
  pl.splits = (int)((kblocks + kb_per - 1) / kb_per);
  int64_t total = tiles * pl.splits;
  pl.grid = (int)(total < num_sms ? total : num_sms);
  return pl;
}

// defined in gemm_tc_64.cu / gemm_tc_128.cu / gemm_tc_256.cu (one translation unit per tile width)
cudaError_t tc_launch_64(bool a_mn, bool b_mn, const TcParams& p, int grid, cudaStream_t st);
cudaError_t tc_launch_128(bool a_mn, bool b_mn, const TcParams& p, int grid, cudaStream_t st);
cudaError_t tc_launch_256(bool a_mn, bool b_mn, const TcParams& p, int grid, cudaStream_t st);

// C (or split-K slices in `splitk_ws`) = opA(A) opB(B).  Returns the number of loss partials written
// (= grid) through *n_partials.  When pl.splits > 1 the caller reduces the slices afterwards.
inline cudaError_t launch_gemm_tc(bool ta, bool tb, const GemmArgs& g, const TcPlan& pl, float* splitk_ws,
                                  cudaStream_t st) {
  TcParams p;
  const bool a_mn = ta, b_mn = !tb;
  bool ok = true;
  static int use3d = -1;
  if (use3d < 0) { const char* ev = getenv("MMAE_TMA_3D"); use3d = (ev && ev[0] == '0') ? 0 : 1; }
  p.a3d = (use3d && a_mn && (g.M % 32) == 0) ? 1 : 0;
  p.b3d = (use3d && b_mn && (g.N % 32) == 0) ? 1 : 0;
  if (!a_mn) ok = ok && make_tmap(&p.tmA, g.A, g.M, g.K, g.lda, TC_BK, TC_BM);
  else if (p.a3d) ok = ok && make_tmap_mn3d(&p.tmA, g.A, g.K, g.M, g.lda, TC_BK, TC_BM / 32);
  else       ok = ok && make_tmap(&p.tmA, g.A, g.K, g.M, g.lda, 32, TC_BK, true);
  if (!b_mn) ok = ok && make_tmap(&p.tmB, g.B, g.N, g.K, g.ldb, TC_BK, pl.bn);
  else if (p.b3d) ok = ok && make_tmap_mn3d(&p.tmB, g.B, g.K, g.N, g.ldb, TC_BK, pl.bn / 32);
  else       ok = ok && make_tmap(&p.tmB, g.B, g.K, g.N, g.ldb, 32, TC_BK, true);
  if (!ok) return cudaErrorInvalidValue;
  p.M = g.M; p.N = g.N; p.K = g.K;
  p.m_blocks = pl.m_blocks; p.n_blocks = pl.n_blocks; p.splits = pl.splits; p.k_per_split = pl.k_per_split;
  p.ep = g.ep;
  if (pl.splits > 1) {
    p.C = splitk_ws; p.ldc = g.N; p.split_stride = g.M * g.N;
    p.ep.mode = EPI_PLAIN; p.ep.beta = 0.f; p.ep.loss_partials = nullptr;
  } else {
    p.C = g.C; p.ldc = g.ldc; p.split_stride = 0;
  }
  if (pl.bn == 256) return tc_launch_256(a_mn, b_mn, p, pl.grid, st);
  if (pl.bn == 128) return tc_launch_128(a_mn, b_mn, p, pl.grid, st);
  return tc_launch_64(a_mn, b_mn, p, pl.grid, st);
}

// ------------------------------------------------------------------ two-SM variant (gemm_tc2.cu): 256 x 256 tiles on CTA pairs
cudaError_t tc2_launch(bool a_mn, bool b_mn, bool noise, const TcParams& p, int grid, cudaStream_t st);

inline bool tc2_eligible(bool ta, bool tb, const GemmArgs& g) {
  static int on = -1;
  if (on < 0) { const char* ev = getenv("MMAE_TC2"); on = (ev && ev[0] == '0') ? 0 : 1; }
  if (!on || !tc_gemm_eligible(ta, tb, g, true)) return false;
  if (g.M < 256 || g.N <= 128) return false;                 // narrower problems keep the one-SM tiles
  if (g.noise.enabled && ta != !tb) return false;            // noise on A: forward (K-major A, K-major B) or its wgrad (both MN-major)
  if (ta && (g.M % 32)) return false;                        // MN-major operands go through the 3-D maps
  if (!tb && (g.N % 32)) return false;
  return true;
}

inline TcPlan tc2_plan(const GemmArgs& g, int num_sms, int max_splits) {
  TcPlan pl;
  const int pairs = num_sms / 2;
  pl.bn = 256;
  pl.m_blocks = (int)((g.M + 255) / 256);
  pl.n_blocks = (int)((g.N + 255) / 256);
  const int64_t tiles = (int64_t)pl.m_blocks * pl.n_blocks;
  const int64_t kblocks = (g.K + TC_BK - 1) / TC_BK;
  int splits = 1;
  if (max_splits > 1) {
    double best = 0.0;
    for (int s = 1; s <= max_splits; ++s) {
      if (s > 1 && kblocks / s < (tiles * s <= pairs ? 8 : 32)) break;
      const int64_t total = tiles * s;
      const int64_t waves = (total + pairs - 1) / pairs;
      const double eff = (double)total / (double)(waves * pairs);
      if (eff > best + 0.03) { best = eff; splits = s; }
    }
  }
  const int64_t kb_per = (kblocks + splits - 1) / splits;
  pl.k_per_split = kb_per * TC_BK;
  pl.splits = (int)((kblocks + kb_per - 1) / kb_per);
  const int64_t total = tiles * pl.splits;
  pl.grid = 2 * (int)(total < pairs ? total : pairs);
  return pl;
}

inline cudaError_t launch_gemm_tc2(bool ta, bool tb, const GemmArgs& g, const TcPlan& pl, float* splitk_ws, cudaStream_t st) {
  TcParams p;
  const bool a_mn = ta, b_mn = !tb;
  bool ok = true;
  p.a3d = a_mn ? 1 : 0; p.b3d = b_mn ? 1 : 0;
  if (!a_mn) ok = ok && make_tmap(&p.tmA, g.A, g.M, g.K, g.lda, TC_BK, TC_BM);
  else ok = ok && make_tmap_mn3d(&p.tmA, g.A, g.K, g.M, g.lda, TC_BK, TC_BM / 32);
  if (!b_mn) ok = ok && make_tmap(&p.tmB, g.B, g.N, g.K, g.ldb, TC_BK, 128);            // each CTA of the pair stages half of the 256 columns
  else ok = ok && make_tmap_mn3d(&p.tmB, g.B, g.K, g.N, g.ldb, TC_BK, 128 / 32);
  if (!ok) return cudaErrorInvalidValue;
  p.M = g.M; p.N = g.N; p.K = g.K;
  p.m_blocks = pl.m_blocks; p.n_blocks = pl.n_blocks; p.splits = pl.splits; p.k_per_split = pl.k_per_split;
  p.ep = g.ep;
  if (pl.splits > 1) {
    p.C = splitk_ws; p.ldc = g.N; p.split_stride = g.M * g.N;
    p.ep.mode = EPI_PLAIN; p.ep.beta = 0.f; p.ep.loss_partials = nullptr;
  } else {
    p.C = g.C; p.ldc = g.ldc; p.split_stride = 0;
  }
  p.nz = g.noise; p.nz_aligned = g.noise_aligned32;
  // Row-layout epilogue: thread = accumulator row, 16-byte shared-memory stores into a 128B-swizzled tile, one TMA store per
  // 32 x 32 chunk, auxiliary operand (loss target / saved activation) read as 16-byte global loads issued before the
  // accumulator is awaited.  A third of the memory instructions of the column-per-lane epilogue it replaces, which
  // bounded the dgrad and the K = 256 launches of the wide step.
  static int tma_epi_on = -1;
  if (tma_epi_on < 0) { const char* ev = getenv("MMAE_TMA_EPI"); tma_epi_on = (ev && ev[0] == '0') ? 0 : 1; }
  p.tma_epi = 0;
  if (tma_epi_on && !a_mn && !b_mn && !g.noise.enabled && pl.splits == 1 && p.ep.mode != EPI_PLAIN && p.ep.beta == 0.f &&
      p.ep.keep >= 1.f && (g.N & 3) == 0) {
    int64_t ldaux = 0; const float* auxp = nullptr;
    if (p.ep.mode == EPI_LOSS_TRAIN || p.ep.mode == EPI_LOSS_PRED) { auxp = p.ep.target; ldaux = p.ep.ldt; }
    else if (p.ep.mode == EPI_DGRAD) { auxp = p.ep.saved; ldaux = p.ep.lds; }
    const bool aux_ok = !auxp || ((ldaux & 3) == 0 && (reinterpret_cast<uintptr_t>(auxp) & 15) == 0);
    if (aux_ok && make_tmap_io(&p.tmC, p.C, g.M, g.N, p.ldc)) p.tma_epi = 1;
  }
  if (p.ep.aux_rows && !p.tma_epi) return cudaErrorInvalidValue;      // only the row-layout epilogue gathers its target rows
  return tc2_launch(a_mn, b_mn, g.noise.enabled != 0, p, pl.grid, st);
}

}  // namespace mmae
