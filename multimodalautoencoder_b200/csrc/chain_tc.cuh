// Whole-network tcgen05 kernel for the small-width MMAE configs (north star: "a persistent whole-network
// kernel ... streams batches from HBM").  One launch runs a CHAIN of dense layers over 128-row tiles:
//
//   op 0     A = X tile streamed from HBM by TMA (k-chunks of 32 through a 4- to 6-deep ring, so the rows of the
//            NEXT tile are already in flight while this tile computes), B = W_0 (K-major) k-chunks -> TMEM
//   op i>0   A = the activated output of op i-1, which never leaves the SM: the epilogue of op i-1 reads
//            its accumulator with tcgen05.ld (one thread = one row, 32 columns), applies bias + activation
//            (+ dropout), rounds to tf32 and writes it back IN PLACE with tcgen05.st; op i's tcgen05.mma takes
//            A straight from TMEM ("TS" form) and starts on k-chunk c as soon as the 32 columns of chunk c
//            are written (per-chunk mbarriers), B = W_i k-chunks from the L2-resident weights via a TMA ring
//   last op  bias + loss (sigmoid-CE / RMSE / CE partial sums) + delta_L or decoded_X, in the same row layout
//
// All global traffic of the epilogues goes through TMA: outputs are written into 128B-swizzled [32 x 32]
// smem tiles (conflict-free 16-byte accesses, one row per thread) and leave with cp.async.bulk.tensor
// stores; the loss target arrives the same way, one chunk ahead.  Per sample the kernel reads X once
// (twice when a loss / target is requested) and writes only what the caller asked for (reconstruction; in
// training also the saved activations the backward GEMMs need).  encode (:454-475) + decode (:499-518) +
// loss (:381-390) of /root/reference/multimodal_autoencoder.py in one launch.
//
// TMEM plan (512 columns): the last op's accumulator has its own region (so its epilogue -- the longest --
// overlaps the first MMAs of the next tile); the other ops alternate between two regions R0 / R1 (op i+1
// reads its A from the region op i wrote, op i+2 may overwrite it: tcgen05.mma executes in issue order).
#pragma once
#include <vector>

#include "gemm_tc.cuh"

namespace mmae {

constexpr int CH_MAX_OPS = 8;
constexpr int CH_MAX_CHUNKS = 8;         // 32-column chunks of a non-final op (N <= 256)
// 20 warps: X producer, W producer, MMA issuer, TMEM owner; 8 "hidden" epilogue warps (non-final ops: activations
// back into TMEM) and 8 "output" epilogue warps (final op: loss + result), so that the output epilogue of tile t runs
// concurrently with the whole hidden chain of tile t+1.
constexpr int CH_THREADS = 640;
constexpr int CH_HID_WARP0 = 4;
constexpr int CH_OUT_WARP0 = 12;
constexpr int CH_EPI_WARPS = 8;                  // per group
constexpr int CH_MAX_XSTAGES = 8;
constexpr int CH_XBYTES = TC_BM * TC_BK * 4;     // 16 KB: [128 rows][32 k]
// X ring + W ring (+ the hidden warps' staging tiles when saved activations are wanted) share one pool; the split is
// chosen per launch (weight k-chunks: [n_chunk n][32 k] slots; X k-chunks: 16 KB slots)
constexpr int CH_POOL_BYTES = 156 * 1024;
constexpr int CH_MAX_WSTAGES = 8;
constexpr int CH_EPI_TILE = 32 * 32 * 4;         // 4 KB: [32 rows][32 cols], 128B swizzle
constexpr int CH_HID_BYTES = CH_EPI_WARPS * CH_EPI_TILE;       // one staging tile per hidden warp (inside the pool)
constexpr int CH_OUT_BYTES = CH_EPI_WARPS * 2 * CH_EPI_TILE;   // two per output warp
constexpr int CH_BIAS_FLOATS = 1024;
constexpr int CH_BAR_BYTES = 1024;
constexpr int CH_SMEM = CH_POOL_BYTES + CH_OUT_BYTES + CH_BIAS_FLOATS * 4 + CH_BAR_BYTES + 1024;
constexpr int CH_TMEM_COLS = 512;
static_assert(CH_SMEM <= 232448, "chain kernel shared memory exceeds the 227 KB per-CTA limit");

struct ChainOp {
  int a_tmem;          // 0: A streamed from global (op 0 only); 1: A = TMEM columns [a_col, a_col + K)
  int a_col;
  int K;               // contraction length
  int N;               // real output width
  int n_chunk;         // accumulator columns per MMA (multiple of 32, <= 256)
  int n_chunks;
  int d_col;           // accumulator column of chunk 0
  int writeback;       // epilogue = bias + act (+dropout), written back to TMEM as the next op's A operand
  int bias_off;        // offset of this op's (zero-padded) bias in the shared bias table
  int has_out;         // a global copy of the output is wanted (tmO valid)
  Epilogue ep;
};

struct ChainParams {
  CUtensorMap tmA;
  CUtensorMap tmB[CH_MAX_OPS];
  CUtensorMap tmO[CH_MAX_OPS];     // outputs, [32 x 32] boxes, plain fp32
  CUtensorMap tmT;                 // loss target of the last op
  ChainOp op[CH_MAX_OPS];
  int nops;
  int w_slot_bytes, w_stages, x_stages, hid_tiles;
  int64_t M;
  int m_tiles;
  // fill-in: detect the missing blocks of each row (sum == -width, data_funcs.py:366-381) from the X tile while it sits in
  // the shared-memory ring, instead of a separate pass over X.  starts = modality start columns [num_mod + 1] (device).
  int scan_miss, num_mod; const int32_t* starts;
  unsigned stagger_ns; // per-CTA start offset step: CTAs that all start together also all load / compute / store together,
                       // so HBM and L2 see bursts; offsetting them over one tile period smooths the demand (0 = off)
  long long* trace;    // debug: clock64 stamps of CTA 0 ([tile it][op][4]: mma start, mma issued, epi start, epi end), or null
};

// One layer of the chain as the engine describes it: y = epilogue(a . W), W given K-major ([N rows, K cols]).
struct ChainLayer {
  const float* Wkm; int64_t ldw; int K, N;
  Epilogue ep; float* out; int64_t ldo;
};

inline int ch_round_up(int x, int m) { return (x + m - 1) / m * m; }

// Fills `p` for X[M, K0] (row stride ldx) pushed through `layers`.  Returns false when the chain does not fit the
// kernel (TMEM columns, alignment); the caller then uses the per-layer GEMMs.
inline bool chain_build(ChainParams& p, const float* X, int64_t M, int64_t ldx, const std::vector<ChainLayer>& layers) {
  auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  const int n = (int)layers.size();
  if (n < 2 || n > CH_MAX_OPS || M < 1) return false;
  if ((ldx & 3) || !al(X)) return false;
  int r0 = 0, r1 = 0;
  for (int i = 0; i < n; ++i) {
    const ChainLayer& l = layers[i];
    if ((l.ldw & 3) || !al(l.Wkm) || l.K < 8 || l.N < 1) return false;
    if (l.out && ((l.ldo & 3) || !al(l.out))) return false;
    if (i > 0 && l.K != layers[i - 1].N) return false;
    if (i < n - 1) {
      const int np = ch_round_up(l.N, 32);
      if (np > 256) return false;
      if (i & 1) r1 = std::max(r1, np); else r0 = std::max(r0, np);
    }
  }
  const ChainLayer& last = layers[n - 1];
  if (!last.out) return false;
  if (last.ep.target && ((last.ep.ldt & 3) || !al(last.ep.target))) return false;
  const int np_last = ch_round_up(last.N, 32);
  const int nch = (np_last + 255) / 256;
  const int n_chunk_last = ch_round_up((np_last + nch - 1) / nch, 32);
  const int w_last = nch * n_chunk_last;
  if (w_last + r0 + r1 > CH_TMEM_COLS) return false;
  const int col_r0 = w_last, col_r1 = w_last + r0;
  memset(&p, 0, sizeof(p));
  p.nops = n; p.M = M; p.m_tiles = (int)((M + TC_BM - 1) / TC_BM);
  if (!make_tmap(&p.tmA, X, M, layers[0].K, ldx, TC_BK, TC_BM)) return false;
  int bias_off = 0, max_chunk = 0;
  for (int i = 0; i < n; ++i) {
    const ChainLayer& l = layers[i];
    ChainOp& o = p.op[i];
    const bool is_last = i == n - 1;
    o.a_tmem = i > 0; o.K = l.K; o.N = l.N;
    o.n_chunk = is_last ? n_chunk_last : ch_round_up(l.N, 32);
    o.n_chunks = is_last ? nch : 1;
    o.d_col = is_last ? 0 : ((i & 1) ? col_r1 : col_r0);
    o.a_col = i > 0 ? p.op[i - 1].d_col : 0;
    o.writeback = is_last ? 0 : 1;
    if (o.writeback && l.ep.mode != EPI_BIAS_ACT && l.ep.mode != EPI_DGRAD) return false;
    if (is_last && l.ep.mode != EPI_LOSS_TRAIN && l.ep.mode != EPI_LOSS_PRED && l.ep.mode != EPI_DGRAD) return false;
    if ((l.ep.mode == EPI_DGRAD) != (layers[0].ep.mode == EPI_DGRAD)) return false;      // one direction per launch
    if (l.ep.mode == EPI_DGRAD) {        // backward chain: the saved activation is the auxiliary operand of the epilogue
      if (l.ep.beta != 0.f || (l.N & 3)) return false;
      if (l.ep.saved && ((l.ep.lds & 3) || !al(l.ep.saved))) return false;
      if (is_last && (!l.ep.saved || l.ep.target != l.ep.saved || l.ep.ldt != l.ep.lds)) return false;   // final op: aux tile by TMA
    }
    o.bias_off = bias_off; bias_off += o.n_chunk * o.n_chunks;
    o.has_out = l.out ? 1 : 0; o.ep = l.ep;
    max_chunk = std::max(max_chunk, o.n_chunk);
    if (!make_tmap(&p.tmB[i], l.Wkm, l.N, l.K, l.ldw, TC_BK, o.n_chunk)) return false;
    if (l.out && !make_tmap_io(&p.tmO[i], l.out, M, l.N, l.ldo)) return false;
  }
  if (bias_off > CH_BIAS_FLOATS - 128) return false;      // the last 512 bytes hold the fill-in column -> modality table
  if (last.ep.fill_bits && bias_off > CH_BIAS_FLOATS - 128 - 256) return false;      // 1 KB before the table: per-row missing-block bits
  if (last.ep.fill_bits && (last.N > 512 || !last.ep.target || last.ep.mode != EPI_LOSS_PRED)) return false;
  if (last.ep.target && !make_tmap_io(&p.tmT, last.ep.target, M, last.N, last.ep.ldt)) return false;
  p.w_slot_bytes = ch_round_up(max_chunk * TC_BK * 4, 1024);
  p.hid_tiles = 0;
  for (int i = 0; i + 1 < n; ++i) if (layers[i].out) p.hid_tiles = 1;
  // pool split: 3 weight slots (L2 latency), the rest to X so that the next tile's rows stream in during this tile
  const int pool = CH_POOL_BYTES - (p.hid_tiles ? CH_HID_BYTES : 0);
  static const int env_ws = getenv("MMAE_CHAIN_WS") ? atoi(getenv("MMAE_CHAIN_WS")) : 3;
  p.w_stages = std::min(CH_MAX_WSTAGES, std::max(2, env_ws));
  while (p.w_stages > 2 && pool - p.w_stages * p.w_slot_bytes < 2 * CH_XBYTES) --p.w_stages;
  p.x_stages = std::min(CH_MAX_XSTAGES, (pool - p.w_stages * p.w_slot_bytes) / CH_XBYTES);
  if (p.w_stages < 2 || p.x_stages < 2) return false;
  return true;
}

// defined in chain_tc.cu
cudaError_t chain_launch(const ChainParams& p, int grid, cudaStream_t st);

}  // namespace mmae
