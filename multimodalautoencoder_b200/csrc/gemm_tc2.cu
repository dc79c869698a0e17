// Two-SM tcgen05 GEMM (cta_group::2): a CTA pair (cluster of 2, same TPC) computes one 256 x 256 tile.
//
// Why: kind::tf32 operands are 4 bytes wide.  A single-SM 128 x 256 tile reads A (16 KB) + B (32 KB) per 32-wide
// k-block from shared memory while the MMAs of that k-block take 256 clocks: 192 B/clk against the 128 B/clk an SM's
// shared memory delivers (ncu: tensor pipe active 80 % at best).  With cta_group::2 each CTA stages its own 128 rows
// of A and only HALF of B (128 of the 256 columns); the tensor cores of both SMs read both halves.  32 KB per
// k-block per SM = 128 B/clk, and the smaller stage allows a 6-deep ring.
//
//   CTA rank r of the pair:  A rows [256*mb + 128*r, +128),  B columns [256*nb + 128*r, +128)
//   TMA (cp.async.bulk.tensor ... cta_group::2): both CTAs' loads complete_tx on the LEADER's full barrier
//   MMA: one lane of the leader CTA issues tcgen05.mma.cta_group::2 (M = 256, N = 256, K = 8);
//        tcgen05.commit ... multicast::cluster frees the stage / publishes the accumulator in BOTH CTAs
//   TMEM: 2 accumulator stages x 256 columns in each CTA (its 128 rows); each CTA's 8 epilogue warps drain their
//        own rows with the same fused epilogues as the one-SM kernel and arrive on the leader's "drained" barrier.
#include <cooperative_groups.h>

#include "gemm_tc_kernel.cuh"

namespace mmae {

constexpr int TC2_BN = 256;
constexpr int TC2_BN_HALF = 128;
constexpr int TC2_STAGES = 6;
constexpr int TC2_ABYTES = TC_BM * TC_BK * 4;            // 16 KB
constexpr int TC2_BBYTES = TC2_BN_HALF * TC_BK * 4;      // 16 KB
constexpr int TC2_STAGE_BYTES = TC2_ABYTES + TC2_BBYTES;
constexpr int TC2_EPI_BYTES = TC_EPI_WARPS * 32 * TC_STAGE_LD * 4;
constexpr int TC2_BAR_BYTES = 512;
constexpr int TC2_SMEM = TC2_STAGES * TC2_STAGE_BYTES + TC2_EPI_BYTES + 1024 + TC2_BAR_BYTES;
static_assert(TC2_EPI_BYTES % 1024 == 0, "barriers follow the staging region");
static_assert(TC2_SMEM <= 232448, "two-SM GEMM shared memory exceeds the per-CTA limit");
constexpr int TC2_THREADS_NOISE = TC_THREADS + 128;      // + 4 warps that apply the mask / noise to the A tile in shared memory
constexpr int TC2_PATCH_WARP0 = TC_THREADS / 32;
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;              // shared::cluster address of the same offset in the pair's leader CTA

__device__ __forceinline__ void tma2_load_2d(const CUtensorMap* tm, uint32_t leader_bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(tm), "r"(leader_bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma2_load_3d(const CUtensorMap* tm, uint32_t leader_bar, void* dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(tm), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tc2_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on the barrier at this offset in BOTH CTAs of the pair once the MMAs issued so far have retired
__device__ __forceinline__ void tc2_commit_both(uint32_t bar_saddr) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar_saddr), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerMask) : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader_release(uint32_t bar_saddr) {      // makes this CTA's shared-memory writes visible to the pair
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_saddr & kPeerMask) : "memory");
}
__device__ __forceinline__ void mbar_wait_acquire_cluster(uint32_t bar_saddr, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar_saddr), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void sts32(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }

// Block-mask + zero noise (multimodal_autoencoder.py:668-702) on 32 consecutive stored columns of one batch row that sit
// in shared memory at byte offsets base + ((4*j) ^ sw), j = 0..31 (sw: the row's swizzle term).  Block mask wins (:695
// after :683).  `feat0` = first feature column of the 32.
__device__ __forceinline__ void patch_row32(const NoiseView& nz, int aligned, uint32_t base, uint32_t sw, uint32_t zb, uint32_t mbits, int feat0, int nfeat) {
  if (aligned) {
    if ((mbits >> __ldg(nz.col_mod + feat0)) & 1u) {
#pragma unroll
      for (int j = 0; j < 32; ++j) sts32(base + ((uint32_t)(4 * j) ^ sw), nz.mask_with);
    } else {
      while (zb) { const int j = __ffs(zb) - 1; zb &= zb - 1; sts32(base + ((uint32_t)(4 * j) ^ sw), 0.f); }
    }
  } else {
    for (int j = 0; j < 32 && feat0 + j < nfeat; ++j) {
      if ((mbits >> __ldg(nz.col_mod + feat0 + j)) & 1u) sts32(base + ((uint32_t)(4 * j) ^ sw), nz.mask_with);
      else if ((zb >> j) & 1u) sts32(base + ((uint32_t)(4 * j) ^ sw), 0.f);
    }
  }
}


// ------------------------------------------------------------------ row-layout epilogue (thread = accumulator row)
__device__ __forceinline__ uint32_t rt_chunk(uint32_t tile_saddr, int lane, int q) { return tile_saddr + lane * 128 + ((q ^ (lane & 7)) << 4); }
__device__ __forceinline__ void rt_sts128(uint32_t a, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void rt_tma_store(const CUtensorMap* tm, uint32_t src_saddr, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tm), "r"(src_saddr), "r"(c0), "r"(c1) : "memory");
}
struct RtAux { float4 v[8]; };
// Auxiliary values (loss target / saved activation) of the 32 x 32 chunk at (row0, col0), COALESCED: load i of lane l
// covers row row0 + l/8 + 4i, columns col0 + 4 (l%8) .. +3, so every instruction reads four full 128-byte row segments.
// The values reach the row-per-thread layout through the warp's shared-memory tile (rt_stage_aux).
__device__ __forceinline__ void rt_load_aux(const float* auxp, int64_t ldaux, int64_t row0, int64_t M, int64_t col0, int64_t N, int lane, RtAux& ax,
                                            const int64_t* aux_rows = nullptr) {
  const int q = lane & 7;
  const bool col_ok = auxp != nullptr && col0 + q * 4 < N;
  const float* sp = auxp + col0 + q * 4;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t row = row0 + (lane >> 3) + 4 * i;
    if (col_ok && row < M) {
      const int64_t arow = aux_rows ? __ldg(aux_rows + row) : row;        // gathered target: the row of the dataset this batch row came from
      asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                   : "=f"(ax.v[i].x), "=f"(ax.v[i].y), "=f"(ax.v[i].z), "=f"(ax.v[i].w) : "l"(sp + arow * ldaux));
    } else ax.v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
__device__ __forceinline__ float4 rt_lds128(uint32_t a) {
  float4 v; asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a)); return v;
}
// coalesced-load layout -> this thread's row (through the 128B-swizzled tile; conflict-free both ways)
__device__ __forceinline__ void rt_stage_aux(uint32_t tile_s, int lane, RtAux& ax) {
  const int q = lane & 7;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int rr = (lane >> 3) + 4 * i;
    rt_sts128(tile_s + rr * 128 + ((q ^ (rr & 7)) << 4), ax.v[i]);
  }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 8; ++j) ax.v[j] = rt_lds128(rt_chunk(tile_s, lane, j));
}
// column sums over the 32 rows (lanes) of a chunk held one row per lane: transposing butterfly, 31 shuffles, fixed
// summation tree; lane l returns the sum of column l
__device__ __forceinline__ float rt_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int j = 0; j < 16; ++j) { const bool up = lane & 16; const float send = up ? v[j] : v[j + 16]; const float keep = up ? v[j + 16] : v[j]; v[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16); }
#pragma unroll
  for (int j = 0; j < 8; ++j) { const bool up = lane & 8; const float send = up ? v[j] : v[j + 8]; const float keep = up ? v[j + 8] : v[j]; v[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8); }
#pragma unroll
  for (int j = 0; j < 4; ++j) { const bool up = lane & 4; const float send = up ? v[j] : v[j + 4]; const float keep = up ? v[j + 4] : v[j]; v[j] = keep + __shfl_xor_sync(0xffffffffu, send, 4); }
#pragma unroll
  for (int j = 0; j < 2; ++j) { const bool up = lane & 2; const float send = up ? v[j] : v[j + 2]; const float keep = up ? v[j + 2] : v[j]; v[j] = keep + __shfl_xor_sync(0xffffffffu, send, 2); }
  { const bool up = lane & 1; const float send = up ? v[0] : v[1]; const float keep = up ? v[1] : v[0]; v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 1); }
  return v[0];
}

// One 32 x 32 chunk: r = this row's accumulators.  MODE / SUB / DROP as in epi_rows (gemm_tc_kernel.cuh).
template <int MODE, int SUB, bool DROP>
__device__ __forceinline__ void rt_chunk_apply(const Epilogue& ep, const uint32_t (&r)[32], const RtAux& ax, int64_t col0, int64_t N,
                                               int64_t grow, bool row_valid, uint32_t tile_s, int lane, float& loss_acc, float* colsum_row) {
  constexpr bool kLoss = (MODE == EPI_LOSS_TRAIN || MODE == EPI_LOSS_PRED);
  const bool has_t = kLoss && ep.target != nullptr;
  float outv[32];
  float lsum = 0.f;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool col_ok = col0 + q * 4 < N;
    if ((MODE == EPI_BIAS_ACT || kLoss) && ep.bias && col_ok) b4 = __ldg(reinterpret_cast<const float4*>(ep.bias + col0) + q);     // warp-uniform: broadcast
    const float bs[4] = {b4.x, b4.y, b4.z, b4.w};
    const float as[4] = {ax.v[q].x, ax.v[q].y, ax.v[q].z, ax.v[q].w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float acc = __uint_as_float(r[q * 4 + e]);
      float out;
      if (MODE == EPI_BIAS_ACT) {
        out = act_fast_t<SUB>(acc + bs[e]);
        if (DROP) {
          uint32_t w = philox_word((uint64_t)grow * (uint64_t)ep.drop_width + (uint64_t)(col0 + q * 4 + e), ep.drop_stream, __ldg(ep.step), ep.seed);
          out = ((w >> 8) < ep.keep_thr) ? out / ep.keep : 0.f;
        }
      } else if (MODE == EPI_DGRAD) {
        float g = acc, h = as[e];
        if (DROP) {
          uint32_t w = philox_word((uint64_t)grow * (uint64_t)ep.drop_width + (uint64_t)(col0 + q * 4 + e), ep.drop_stream, __ldg(ep.step), ep.seed);
          if ((w >> 8) < ep.keep_thr) { g = g / ep.keep; h = h * ep.keep; } else { g = 0.f; }
        }
        out = g * dact_t<SUB>(h);
      } else {
        const float l = acc + bs[e], x = as[e];
        float lv;
        if (SUB == MMAE_LOSS_SIGMOID_CE) {
          const float ee = __expf(-fabsf(l));
          const float inv = __fdividef(1.f, 1.f + ee);
          const float sg = l >= 0.f ? inv : ee * inv;
          lv = fmaxf(l, 0.f) - l * x + __logf(1.f + ee);
          out = (MODE == EPI_LOSS_TRAIN) ? (sg - x) : sg;
        } else if (SUB == MMAE_LOSS_RMSE) {
          const float d = l - x;
          lv = d * d;
          out = (MODE == EPI_LOSS_TRAIN) ? d : l;
        } else {
          lv = -x * __logf(l);
          out = (MODE == EPI_LOSS_TRAIN) ? __fdividef(-x, l) : l;
        }
        if (has_t && col_ok) lsum += lv;
      }
      outv[q * 4 + e] = out;
    }
    rt_sts128(rt_chunk(tile_s, lane, q), make_float4(outv[q * 4], outv[q * 4 + 1], outv[q * 4 + 2], outv[q * 4 + 3]));
  }
  if (has_t && row_valid) loss_acc += lsum;
  if (colsum_row) {            // bias gradient: column sums of the stored values over this chunk's rows inside M
    if (!row_valid) {
#pragma unroll
      for (int j = 0; j < 32; ++j) outv[j] = 0.f;
    }
    const float cs = rt_colsum32(outv, lane);
    if (col0 + lane < N) colsum_row[col0 + lane] = cs;
  }
}
template <int MODE, bool DROP>
__device__ __forceinline__ void rt_dispatch_act(const Epilogue& ep, const uint32_t (&r)[32], const RtAux& ax, int64_t col0, int64_t N, int64_t grow,
                                                bool row_valid, uint32_t tile_s, int lane, float& loss_acc, float* colsum_row) {
  switch (ep.act) {
    case MMAE_ACT_RELU: rt_chunk_apply<MODE, MMAE_ACT_RELU, DROP>(ep, r, ax, col0, N, grow, row_valid, tile_s, lane, loss_acc, colsum_row); break;
    case MMAE_ACT_TANH: rt_chunk_apply<MODE, MMAE_ACT_TANH, DROP>(ep, r, ax, col0, N, grow, row_valid, tile_s, lane, loss_acc, colsum_row); break;
    case MMAE_ACT_SOFTSIGN: rt_chunk_apply<MODE, MMAE_ACT_SOFTSIGN, DROP>(ep, r, ax, col0, N, grow, row_valid, tile_s, lane, loss_acc, colsum_row); break;
    case MMAE_ACT_SOFTPLUS: rt_chunk_apply<MODE, MMAE_ACT_SOFTPLUS, DROP>(ep, r, ax, col0, N, grow, row_valid, tile_s, lane, loss_acc, colsum_row); break;
    default: rt_chunk_apply<MODE, MMAE_ACT_LINEAR, DROP>(ep, r, ax, col0, N, grow, row_valid, tile_s, lane, loss_acc, colsum_row); break;
  }
}
template <int MODE>
__device__ __forceinline__ void rt_dispatch_loss(const Epilogue& ep, const uint32_t (&r)[32], const RtAux& ax, int64_t col0, int64_t N, int64_t grow,
                                                 bool row_valid, uint32_t tile_s, int lane, float& loss_acc, float* colsum_row) {
  switch (ep.loss) {
    case MMAE_LOSS_SIGMOID_CE: rt_chunk_apply<MODE, MMAE_LOSS_SIGMOID_CE, false>(ep, r, ax, col0, N, grow, row_valid, tile_s, lane, loss_acc, colsum_row); break;
    case MMAE_LOSS_RMSE: rt_chunk_apply<MODE, MMAE_LOSS_RMSE, false>(ep, r, ax, col0, N, grow, row_valid, tile_s, lane, loss_acc, colsum_row); break;
    default: rt_chunk_apply<MODE, MMAE_LOSS_CE, false>(ep, r, ax, col0, N, grow, row_valid, tile_s, lane, loss_acc, colsum_row); break;
  }
}
__device__ __forceinline__ void rt_dispatch(const Epilogue& ep, const uint32_t (&r)[32], const RtAux& ax, int64_t col0, int64_t N, int64_t grow,
                                            bool row_valid, uint32_t tile_s, int lane, float& loss_acc, float* colsum_row) {
  // (dropout steps keep the column-per-lane epilogue: the host side leaves tma_epi off when keep < 1)
  switch (ep.mode) {
    case EPI_BIAS_ACT: rt_dispatch_act<EPI_BIAS_ACT, false>(ep, r, ax, col0, N, grow, row_valid, tile_s, lane, loss_acc, colsum_row); break;
    case EPI_DGRAD: rt_dispatch_act<EPI_DGRAD, false>(ep, r, ax, col0, N, grow, row_valid, tile_s, lane, loss_acc, colsum_row); break;
    case EPI_LOSS_TRAIN: rt_dispatch_loss<EPI_LOSS_TRAIN>(ep, r, ax, col0, N, grow, row_valid, tile_s, lane, loss_acc, colsum_row); break;
    default: rt_dispatch_loss<EPI_LOSS_PRED>(ep, r, ax, col0, N, grow, row_valid, tile_s, lane, loss_acc, colsum_row); break;
  }
}

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <bool A_MN, bool B_MN, bool NOISE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NOISE ? TC2_THREADS_NOISE : TC_THREADS, 1) gemm_tc2_kernel(const __grid_constant__ TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* epi_smem = smem + TC2_STAGES * TC2_STAGE_BYTES;                  // 1024-byte aligned (TMA-store tiles are 128B-swizzled)
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_smem + TC2_EPI_BYTES);
  uint64_t* full_bar = bars;                          // [TC2_STAGES]  (the leader's is the one that counts)
  uint64_t* empty_bar = bars + TC2_STAGES;            // [TC2_STAGES]
  uint64_t* tfull_bar = bars + 2 * TC2_STAGES;        // [2]
  uint64_t* tempty_bar = bars + 2 * TC2_STAGES + 2;   // [2]  leader's: 16 arrivals (8 epilogue warps x 2 CTAs)
  uint64_t* afull_bar = bars + 2 * TC2_STAGES + 4;    // [TC2_STAGES]  NOISE: this CTA's A tile has landed (local)
  uint64_t* aok_bar = afull_bar + TC2_STAGES;         // [TC2_STAGES]  NOISE: both CTAs' A tiles are patched (leader's: one warp per CTA)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aok_bar + TC2_STAGES);
  float* epi_red = reinterpret_cast<float*>(tmem_slot + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int64_t tiles_mn = (int64_t)p.m_blocks * p.n_blocks;       // m_blocks counts 256-row blocks here
  const int64_t num_tiles = tiles_mn * p.splits;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tmB) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < TC2_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], 2 * TC_EPI_WARPS); }
    for (int s = 0; s < TC2_STAGES; ++s) { mbar_init(&afull_bar[s], 1); mbar_init(&aok_bar[s], 2); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {   // both CTAs of the pair allocate (same logical warp, same slot address)
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();            // the peer's barriers exist before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto decode = [&](int64_t t, int& mb, int& nb, int& sp) {
    sp = (int)(t / tiles_mn);
    int64_t r = t - (int64_t)sp * tiles_mn;
    mb = (int)(r / p.n_blocks);
    nb = (int)(r - (int64_t)mb * p.n_blocks);
  };

  if (warp == 0) {
    // ===================== TMA producer (one lane per CTA): own rows of A, own half of B =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      const uint32_t full_s = smem_u32(full_bar);
      for (int64_t t = pair; t < num_tiles; t += num_pairs) {
        int mb, nb, sp; decode(t, mb, nb, sp);
        const int64_t kb0 = (int64_t)sp * p.k_per_split;
        const int64_t kend = min(p.K, kb0 + p.k_per_split);
        const int m0 = mb * 256 + (int)rank * TC_BM;
        const int n0 = nb * TC2_BN + (int)rank * TC2_BN_HALF;
        for (int64_t k = kb0; k < kend; k += TC_BK) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * TC2_STAGE_BYTES;
          uint8_t* sb = sa + TC2_ABYTES;
          const uint32_t lbar = (full_s + stage * 8u) & kPeerMask;
          if (NOISE) {      // A lands on this CTA's own barrier: the patch warps touch it before the MMA may
            if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * TC2_BBYTES);
            mbar_expect_tx(&afull_bar[stage], TC2_ABYTES);
            if (!A_MN) tma_load_2d(&p.tmA, &afull_bar[stage], sa, (int)k, m0);
            else tma_load_3d(&p.tmA, &afull_bar[stage], sa, 0, (int)k, m0 / 32);
          } else {
          if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * TC2_STAGE_BYTES);      // both CTAs' bytes land on the leader's barrier
          if (!A_MN) tma2_load_2d(&p.tmA, lbar, sa, (int)k, m0);                      // box {32 k, 128 m}
          else tma2_load_3d(&p.tmA, lbar, sa, 0, (int)k, m0 / 32);                    // box {32 m, 32 k, 4 chunks}
          }
          if (!B_MN) tma2_load_2d(&p.tmB, lbar, sb, (int)k, n0);                      // box {32 k, 128 n}
          else tma2_load_3d(&p.tmB, lbar, sb, 0, (int)k, n0 / 32);                    // box {32 n, 32 k, 4 chunks}
          if (++stage == TC2_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: one lane of the leader CTA =====================
    if (lane == 0 && rank == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                             ((uint32_t)(TC2_BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      const uint64_t a_hi = !A_MN ? make_smem_desc(0, 16, 1024) : make_smem_desc(0, 4096, 512, 1);
      const uint64_t b_hi = !B_MN ? make_smem_desc(0, 16, 1024) : make_smem_desc(0, 4096, 512, 1);
      constexpr uint32_t a_step = !A_MN ? (32 >> 4) : (1024 >> 4);
      constexpr uint32_t b_step = !B_MN ? (32 >> 4) : (1024 >> 4);
      const uint32_t smem_s = smem_u32(smem);
      const uint32_t full_s = smem_u32(full_bar), empty_s = smem_u32(empty_bar), tfull_s = smem_u32(tfull_bar);
      int stage = 0; uint32_t phase = 0;
      int64_t it = 0;
      for (int64_t t = pair; t < num_tiles; t += num_pairs, ++it) {
        int mb, nb, sp; decode(t, mb, nb, sp);
        const int64_t kb0 = (int64_t)sp * p.k_per_split;
        const int64_t kend = min(p.K, kb0 + p.k_per_split);
        const int nkb = (int)((kend - kb0 + TC_BK - 1) / TC_BK);
        const int as = (int)(it & 1); const uint32_t aphase = (uint32_t)((it >> 1) & 1);
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * TC2_BN);
        uint32_t accumulate = 0;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait_addr(full_s + stage * 8u, phase);
          if (NOISE) mbar_wait_acquire_cluster(smem_u32(aok_bar) + stage * 8u, phase);
          tc_fence_after();
          const uint32_t sa = smem_s + stage * TC2_STAGE_BYTES;
          const uint64_t adesc = a_hi | (uint64_t)((sa >> 4) & 0x3FFF);
          const uint64_t bdesc = b_hi | (uint64_t)(((sa + TC2_ABYTES) >> 4) & 0x3FFF);
          tc2_mma_tf32(tmem_d, adesc, bdesc, idesc, accumulate);
          tc2_mma_tf32(tmem_d, adesc + a_step, bdesc + b_step, idesc, 1);
          tc2_mma_tf32(tmem_d, adesc + 2 * a_step, bdesc + 2 * b_step, idesc, 1);
          tc2_mma_tf32(tmem_d, adesc + 3 * a_step, bdesc + 3 * b_step, idesc, 1);
          accumulate = 1;
          tc2_commit_both(empty_s + stage * 8u);        // both CTAs' producers may refill the stage
          if (++stage == TC2_STAGES) { stage = 0; phase ^= 1; }
        }
        tc2_commit_both(tfull_s + as * 8u);             // accumulator complete in both CTAs
      }
    }
  } else if (NOISE && warp >= TC2_PATCH_WARP0) {
    // ===================== patch warps (4 per CTA): mask + noise on the A tile, in shared memory =====================
    // "fused modality-block mask plus noise in the A-operand load of the first encoder GEMM": TMA lands the clean rows,
    // these warps overwrite the masked blocks with mask_with and the 5 % noise cells with 0, then hand the tile to the
    // tensor cores.  One work item = 32 stored columns of one batch row (= one word of the zero bitmap); a stage has
    // 128 items.  Making the patched tile visible to the pair (proxy fence + cluster-scope release) costs ~1 us, far
    // more than the 512 clocks a stage lasts, so the four warps take the k-blocks round-robin -- four stages are in
    // flight -- and each warp covers a whole stage (4 items per lane).
    // MEASURED (wide step, same box): 11.14 ms with a materialised noisy X (noise_apply_kernel, 0.35 ms) against
    // 13.85 ms with this path (12.16 ms with cta-scope signalling, which is not formally sufficient across the pair):
    // the ring is 6 stages deep and HBM latency already uses most of that depth, so any latency added between "tile
    // landed" and "tile usable" lowers the stage rate.  The engine therefore keeps materialisation as the default and
    // this path behind MMAE_FUSE_NOISE=1 (parity-tested either way).
    //   K-major A (forward): item t = tile row (batch row m0 + t), columns k..k+31.
    //   MN-major A (wgrad): item t = (chunk t / 32, batch row k + t % 32), feature columns m0 + 32 (t / 32) .. +31.
    const int pw = warp - TC2_PATCH_WARP0;
    const uint32_t smem_s = smem_u32(smem);
    const uint32_t aok_s = smem_u32(aok_bar);
    const int64_t nrows = !A_MN ? p.M : p.K;
    const int nfeat = !A_MN ? (int)p.K : (int)p.M;
    uint32_t c = 0;                                   // k-blocks produced so far (all tiles): stage = c % STAGES
    for (int64_t t = pair; t < num_tiles; t += num_pairs) {
      int mb, nb, sp; decode(t, mb, nb, sp);
      const int64_t kb0 = (int64_t)sp * p.k_per_split;
      const int64_t kend = min(p.K, kb0 + p.k_per_split);
      const int m0 = mb * 256 + (int)rank * TC_BM;
      auto fetch = [&](int64_t k, uint32_t (&zb)[4], uint32_t (&mbits)[4]) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int tid = lane + 32 * j;
          zb[j] = 0u; mbits[j] = 0u;
          const int64_t brow = !A_MN ? (int64_t)m0 + tid : k + (tid & 31);
          const int feat0 = !A_MN ? (int)k : m0 + (tid >> 5) * 32;
          if (k < kend && brow < nrows && feat0 < nfeat) {
            zb[j] = __ldg(p.nz.zero_bits + brow * p.nz.zw + (feat0 >> 5)); mbits[j] = __ldg(p.nz.mod_bits + brow);
          }
        }
      };
      uint32_t zq[4], mq[4], zn[4], mn[4];
      // first k-block of this tile that is mine
      int64_t k = kb0 + (int64_t)((pw - (int)(c & 3)) & 3) * TC_BK;
      uint32_t cc = c + (uint32_t)((pw - (int)(c & 3)) & 3);
      fetch(k, zq, mq);
      for (; k < kend; k += 4 * TC_BK, cc += 4) {
        fetch(k + 4 * TC_BK, zn, mn);
        const uint32_t stage = cc % TC2_STAGES, phase = (cc / TC2_STAGES) & 1;
        mbar_wait(&afull_bar[stage], phase);
        const uint32_t sa = smem_s + stage * TC2_STAGE_BYTES;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int tid = lane + 32 * j;
          const int feat0 = !A_MN ? (int)k : m0 + (tid >> 5) * 32;
          if (zq[j] | mq[j]) {
            if (!A_MN) patch_row32(p.nz, p.nz_aligned, sa + tid * 128, (uint32_t)(tid & 7) << 4, zq[j], mq[j], feat0, nfeat);
            else patch_row32(p.nz, p.nz_aligned, sa + (tid >> 5) * 4096 + (tid & 31) * 128, (uint32_t)(tid & 3) << 5, zq[j], mq[j], feat0, nfeat);
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic writes -> visible to the tensor cores' reads
        __syncwarp();
        if (lane == 0) mbar_arrive_leader_release(aok_s + stage * 8u);
#pragma unroll
        for (int j = 0; j < 4; ++j) { zq[j] = zn[j]; mq[j] = mn[j]; }
      }
      c += (uint32_t)((kend - kb0 + TC_BK - 1) / TC_BK);
    }
  } else if (warp >= TC_EPI_WARP0 && warp < TC_EPI_WARP0 + TC_EPI_WARPS) {
    // ===================== epilogue warps (8 per CTA): this CTA's 128 rows, all 256 columns =====================
    constexpr bool kRowEpi = !A_MN && !B_MN && !NOISE;      // forward / dgrad GEMMs (K-major operands); wgrads are EPI_PLAIN
    const int quad = warp & 3;
    const int half = (warp - TC_EPI_WARP0) >> 2;
    float* stg = reinterpret_cast<float*>(epi_smem) + (warp - TC_EPI_WARP0) * 32 * TC_STAGE_LD;
    const uint32_t tile_s = smem_u32(epi_smem) + (uint32_t)(warp - TC_EPI_WARP0) * 4096u;
    float loss_acc = 0.f;
    int64_t ldaux; const float* auxp = epilogue_aux_ptr(p.ep, &ldaux);
    int64_t it = 0;
    for (int64_t t = pair; t < num_tiles; t += num_pairs, ++it) {
      int mb, nb, sp; decode(t, mb, nb, sp);
      const int as = (int)(it & 1); const uint32_t aphase = (uint32_t)((it >> 1) & 1);
      const int64_t row0 = (int64_t)mb * 256 + (int64_t)rank * TC_BM + quad * 32;
      if (kRowEpi && p.tma_epi) {
        // ---- row-layout epilogue: thread = row; the first chunk's auxiliary loads travel while the MMAs still run
        const int64_t row = row0 + lane;
        const bool row_valid = row < p.M;
        const int64_t ncol0 = (int64_t)nb * TC2_BN;
        RtAux ax;
        const int64_t* arows = (p.ep.mode == EPI_LOSS_TRAIN || p.ep.mode == EPI_LOSS_PRED) ? p.ep.aux_rows : nullptr;
        rt_load_aux(auxp, ldaux, row0, p.M, ncol0 + half * 32, p.N, lane, ax, arows);
        mbar_wait(&tfull_bar[as], aphase);
        tc_fence_after();
#pragma unroll 1
        for (int ch = half; ch < TC2_BN / 32; ch += 2) {
          const int64_t col0 = ncol0 + ch * 32;
          if (col0 >= p.N || row0 >= p.M) break;                        // warp-uniform: the rest of the tile is outside the matrix
          uint32_t r[32];
          tc_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * TC2_BN + ch * 32), r);
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");      // the previous store has read the tile
          __syncwarp();
          if (auxp) rt_stage_aux(tile_s, lane, ax);
          float* cs_row = p.ep.colsum_partials ? p.ep.colsum_partials + (row0 >> 5) * p.N : nullptr;
          rt_dispatch(p.ep, r, ax, col0, p.N, row + p.ep.row0, row_valid, tile_s, lane, loss_acc, cs_row);
          // next chunk's aux, behind this chunk's arithmetic.  (Measured and dropped: issuing these loads BEFORE the arithmetic
          // into a second register set -- 168 registers with spills -- made the aux-carrying launches 1.3-2.2x slower.)
          if (ch + 2 < TC2_BN / 32) rt_load_aux(auxp, ldaux, row0, p.M, col0 + 64, p.N, lane, ax, arows);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) { rt_tma_store(&p.tmC, tile_s, (int)col0, (int)row0); asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(&tempty_bar[as]);
        continue;
      }
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      float* cbase = p.C + (int64_t)sp * p.split_stride;
#pragma unroll 1
      for (int ch = half; ch < TC2_BN / 32; ch += 2) {
        uint32_t r[32];
        tc_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * TC2_BN + ch * 32), r);
#pragma unroll
        for (int j = 0; j < 32; ++j) stg[lane * TC_STAGE_LD + j] = __uint_as_float(r[j]);   // lane = row
        __syncwarp();
        const int64_t col = (int64_t)nb * TC2_BN + ch * 32 + lane;                          // lane = column
        if (col < p.N && row0 < p.M) {
          EpiRowCtx c;
          c.cp = cbase + row0 * p.ldc + col;
          c.ap = auxp ? auxp + row0 * ldaux + col : nullptr;
          c.ldc = p.ldc; c.ldaux = ldaux; c.stg = stg + lane;
          c.nrows = (int)min((int64_t)32, p.M - row0);
          c.bias_v = p.ep.bias ? __ldg(p.ep.bias + col) : 0.f;
          c.beta = p.ep.beta; c.grow0 = row0 + p.ep.row0; c.col = col;
          c.colsum_out = p.ep.colsum_partials ? p.ep.colsum_partials + (row0 >> 5) * p.N + col : nullptr;
          epi_dispatch(p.ep, c, loss_acc);
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&tempty_bar[as]);     // 16 arrivals (both CTAs) free the accumulator stage
    }
    if (kRowEpi && p.tma_epi && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");     // every output tile has landed
    if (p.ep.loss_partials) {
      float w = warp_sum(loss_acc);
      if (lane == 0) epi_red[warp - TC_EPI_WARP0] = w;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (warp == TC_EPI_WARP0 && lane == 0) {
        float sacc = 0.f;
#pragma unroll
        for (int i = 0; i < TC_EPI_WARPS; ++i) sacc += epi_red[i];
        p.ep.loss_partials[blockIdx.x] = sacc;
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();            // nobody frees TMEM / exits while the pair still works on it
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

template <bool A_MN, bool B_MN, bool NOISE>
static cudaError_t tc2_launch_inst(const TcParams& p, int grid, cudaStream_t st) {
  static bool configured_dev[64] = {};        // the attribute is per device: one flag per device ordinal
  int dev_ = 0; cudaGetDevice(&dev_);
  bool& configured = configured_dev[dev_ & 63];
  auto kern = gemm_tc2_kernel<A_MN, B_MN, NOISE>;
  if (!configured || dev_ >= 64) {      // (ordinals past the table are configured on every launch instead of aliasing)
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TC2_SMEM);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  kern<<<grid, NOISE ? TC2_THREADS_NOISE : TC_THREADS, TC2_SMEM, st>>>(p);      // cluster shape comes from __cluster_dims__
  return cudaGetLastError();
}

cudaError_t tc2_launch(bool a_mn, bool b_mn, bool noise, const TcParams& p, int grid, cudaStream_t st) {
  if (noise) {      // only the first encoder layer carries noise: its forward (K-major A and B) and its wgrad (both MN-major)
    if (!a_mn && !b_mn) return tc2_launch_inst<false, false, true>(p, grid, st);
    if (a_mn && b_mn) return tc2_launch_inst<true, true, true>(p, grid, st);
    return cudaErrorInvalidValue;
  }
  if (!a_mn && !b_mn) return tc2_launch_inst<false, false, false>(p, grid, st);
  if (!a_mn && b_mn) return tc2_launch_inst<false, true, false>(p, grid, st);
  if (a_mn && !b_mn) return tc2_launch_inst<true, false, false>(p, grid, st);
  return tc2_launch_inst<true, true, false>(p, grid, st);
}

}  // namespace mmae
