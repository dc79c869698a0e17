// Device side of the tcgen05 GEMM (see gemm_tc.cuh for the design notes).  Included only by the
// per-tile-width translation units gemm_tc_{64,128,256}.cu.
#pragma once
#include "gemm_tc.cuh"

namespace mmae {

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void mbar_wait_addr(uint32_t bar_saddr, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar_saddr), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void tc_commit_addr(uint32_t bar_saddr) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_saddr) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* tm, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* tm, uint64_t* bar, void* dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor (PTX ISA "tcgen05 shared memory descriptor"), SWIZZLE_128B:
//   bits [0,14) start address >> 4, [16,30) leading byte offset >> 4, [32,46) stride byte offset >> 4,
//   [46,48) version = 1, [61,64) layout type: 2 = 128-byte swizzle (K-major operands),
//   1 = 128-byte swizzle with 32-byte atomicity (the only legal layout for MN-major 32-bit operands).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t layout_type = 2) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout_type << 61;
  return d;
}

// Instruction descriptor for kind::tf32, fp32 accumulate:
//   [4,6) D format = 1 (f32); [7,10) A format = 2 (tf32); [10,13) B format = 2; bit 15 A major (1 = MN);
//   bit 16 B major; [17,23) N >> 3; [24,29) M >> 4.
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N, bool a_mn, bool b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}


// ------------------------------------------------------------------ epilogue row loops
// One lane = one output column; the warp walks the (up to 32) rows of its staged chunk.  Specialised on
// mode / activation / loss / dropout so the hot loop is ~10-30 instructions per row instead of a generic
// switch, with pointer-increment addressing and the auxiliary reads (target / saved activation / old C)
// batched 8 rows ahead.
template <int ACT> __device__ __forceinline__ float act_fast_t(float z) {
  if (ACT == MMAE_ACT_RELU) return fmaxf(z, 0.f);
  if (ACT == MMAE_ACT_TANH) { float e = __expf(-2.f * fabsf(z)); float t = __fdividef(1.f - e, 1.f + e); return z >= 0.f ? t : -t; }
  if (ACT == MMAE_ACT_SOFTSIGN) return __fdividef(z, 1.f + fabsf(z));
  if (ACT == MMAE_ACT_SOFTPLUS) return fmaxf(z, 0.f) + __logf(1.f + __expf(-fabsf(z)));
  return z;
}
template <int ACT> __device__ __forceinline__ float dact_t(float a) {
  if (ACT == MMAE_ACT_RELU) return a > 0.f ? 1.f : 0.f;
  if (ACT == MMAE_ACT_TANH) return 1.f - a * a;
  if (ACT == MMAE_ACT_SOFTSIGN) { float t = 1.f - fabsf(a); return t * t; }
  if (ACT == MMAE_ACT_SOFTPLUS) return 1.f - __expf(-a);
  return 1.f;
}

struct EpiRowCtx {
  float* cp;            // &C[row0, col]
  const float* ap;      // &aux[row0, col] or null
  int64_t ldc, ldaux;
  const float* stg;     // staging + lane (row stride TC_STAGE_LD)
  int nrows;            // rows of this chunk inside M (warp-uniform)
  float bias_v, beta;
  int64_t grow0, col;   // global row (for the dropout stream) and column
  float* colsum_out;    // where this lane's column sum over the chunk's rows goes (null = not wanted)
};

// MODE: EpiMode; SUB: activation (BIAS_ACT / DGRAD) or loss (LOSS_*); DROP: dropout enabled
template <int MODE, int SUB, bool DROP>
__device__ __forceinline__ void epi_rows(const Epilogue& ep, const EpiRowCtx& c, float& loss_acc) {
  constexpr bool kAux = (MODE == EPI_LOSS_TRAIN || MODE == EPI_LOSS_PRED || MODE == EPI_DGRAD);
  const bool use_old = (MODE == EPI_PLAIN || MODE == EPI_DGRAD) && c.beta != 0.f;
  const bool has_aux = kAux && c.ap != nullptr;
  float* cp = c.cp; const float* ap = c.ap; const float* sp = c.stg;
  float csum = 0.f;
  constexpr int RB = kAux ? 16 : 8;        // rows per batch: 16 x 128 B per warp in flight for the aux stream
  for (int i0 = 0; i0 < c.nrows; i0 += RB) {
    float aux[RB], old[RB];
    const int n = min(RB, c.nrows - i0);
#pragma unroll
    for (int u = 0; u < RB; ++u) {
      aux[u] = (has_aux && u < n) ? __ldg(ap + (int64_t)u * c.ldaux) : 0.f;
      old[u] = (use_old && u < n) ? cp[(int64_t)u * c.ldc] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < RB; ++u) {
      if (u < n) {
        const float acc = sp[u * TC_STAGE_LD];
        float out;
        if (MODE == EPI_PLAIN) {
          out = acc + c.beta * old[u];
        } else if (MODE == EPI_BIAS_ACT) {
          out = act_fast_t<SUB>(acc + c.bias_v);
          if (DROP) {
            uint32_t w = philox_word((uint64_t)(c.grow0 + i0 + u) * (uint64_t)ep.drop_width + (uint64_t)c.col, ep.drop_stream, __ldg(ep.step), ep.seed);
            out = ((w >> 8) < ep.keep_thr) ? out / ep.keep : 0.f;
          }
        } else if (MODE == EPI_DGRAD) {
          float g = acc + c.beta * old[u];
          float h = aux[u];
          if (DROP) {
            uint32_t w = philox_word((uint64_t)(c.grow0 + i0 + u) * (uint64_t)ep.drop_width + (uint64_t)c.col, ep.drop_stream, __ldg(ep.step), ep.seed);
            if ((w >> 8) < ep.keep_thr) { g = g / ep.keep; h = h * ep.keep; } else { g = 0.f; }
          }
          out = g * dact_t<SUB>(h);
        } else {   // EPI_LOSS_TRAIN / EPI_LOSS_PRED, SUB = loss
          const float l = acc + c.bias_v, x = aux[u];
          if (SUB == MMAE_LOSS_SIGMOID_CE) {
            const float e = __expf(-fabsf(l));
            const float inv = __fdividef(1.f, 1.f + e);
            const float s = l >= 0.f ? inv : e * inv;
            if (has_aux) loss_acc += fmaxf(l, 0.f) - l * x + __logf(1.f + e);
            out = (MODE == EPI_LOSS_TRAIN) ? (s - x) : s;
          } else if (SUB == MMAE_LOSS_RMSE) {
            const float d = l - x;
            if (has_aux) loss_acc += d * d;
            out = (MODE == EPI_LOSS_TRAIN) ? d : l;
          } else {
            if (has_aux) loss_acc += -x * __logf(l);
            out = (MODE == EPI_LOSS_TRAIN) ? __fdividef(-x, l) : l;
          }
        }
        cp[(int64_t)u * c.ldc] = out;
        csum += out;
      }
    }
    cp += RB * c.ldc; sp += RB * TC_STAGE_LD;
    if (kAux) ap += RB * c.ldaux;
  }
  if (c.colsum_out) *c.colsum_out = csum;      // fixed row order -> deterministic
}

template <int MODE, bool DROP>
__device__ __forceinline__ void epi_rows_act(const Epilogue& ep, const EpiRowCtx& c, float& loss_acc) {
  switch (ep.act) {
    case MMAE_ACT_RELU: epi_rows<MODE, MMAE_ACT_RELU, DROP>(ep, c, loss_acc); break;
    case MMAE_ACT_TANH: epi_rows<MODE, MMAE_ACT_TANH, DROP>(ep, c, loss_acc); break;
    case MMAE_ACT_SOFTSIGN: epi_rows<MODE, MMAE_ACT_SOFTSIGN, DROP>(ep, c, loss_acc); break;
    case MMAE_ACT_SOFTPLUS: epi_rows<MODE, MMAE_ACT_SOFTPLUS, DROP>(ep, c, loss_acc); break;
    default: epi_rows<MODE, MMAE_ACT_LINEAR, DROP>(ep, c, loss_acc); break;
  }
}
template <int MODE>
__device__ __forceinline__ void epi_rows_loss(const Epilogue& ep, const EpiRowCtx& c, float& loss_acc) {
  switch (ep.loss) {
    case MMAE_LOSS_SIGMOID_CE: epi_rows<MODE, MMAE_LOSS_SIGMOID_CE, false>(ep, c, loss_acc); break;
    case MMAE_LOSS_RMSE: epi_rows<MODE, MMAE_LOSS_RMSE, false>(ep, c, loss_acc); break;
    default: epi_rows<MODE, MMAE_LOSS_CE, false>(ep, c, loss_acc); break;
  }
}
__device__ __forceinline__ void epi_dispatch(const Epilogue& ep, const EpiRowCtx& c, float& loss_acc) {
  const bool drop = ep.keep < 1.f;
  switch (ep.mode) {
    case EPI_PLAIN: epi_rows<EPI_PLAIN, 0, false>(ep, c, loss_acc); break;
    case EPI_BIAS_ACT: if (drop) epi_rows_act<EPI_BIAS_ACT, true>(ep, c, loss_acc); else epi_rows_act<EPI_BIAS_ACT, false>(ep, c, loss_acc); break;
    case EPI_DGRAD: if (drop) epi_rows_act<EPI_DGRAD, true>(ep, c, loss_acc); else epi_rows_act<EPI_DGRAD, false>(ep, c, loss_acc); break;
    case EPI_LOSS_TRAIN: epi_rows_loss<EPI_LOSS_TRAIN>(ep, c, loss_acc); break;
    default: epi_rows_loss<EPI_LOSS_PRED>(ep, c, loss_acc); break;
  }
}

// ------------------------------------------------------------------ the kernel
template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(TC_THREADS, 1) gemm_tc_kernel(const __grid_constant__ TcParams p) {
  using Cfg = TcCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* full_bar = bars;                          // [kStages]
  uint64_t* empty_bar = bars + Cfg::kStages;          // [kStages]
  uint64_t* tfull_bar = bars + 2 * Cfg::kStages;      // [2]
  uint64_t* tempty_bar = bars + 2 * Cfg::kStages + 2; // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * Cfg::kStages + 4);
  __shared__ float epi_red[TC_EPI_WARPS];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t tiles_mn = (int64_t)p.m_blocks * p.n_blocks;
  const int64_t num_tiles = tiles_mn * p.splits;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tmB) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], TC_EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {   // TMEM allocation (this warp also frees it)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(Cfg::kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // tile -> (m_blk, n_blk, split): n fastest so concurrently resident CTAs share the A row block in L2
  auto decode = [&](int64_t t, int& mb, int& nb, int& sp) {
    sp = (int)(t / tiles_mn);
    int64_t r = t - (int64_t)sp * tiles_mn;
    mb = (int)(r / p.n_blocks);
    nb = (int)(r - (int64_t)mb * p.n_blocks);
  };

  if (warp == 0) {
    // ===================== TMA producer (whole warp loops, lane 0 issues) =====================
    int stage = 0; uint32_t phase = 0;
    for (int64_t t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      int mb, nb, sp; decode(t, mb, nb, sp);
      const int64_t kb0 = (int64_t)sp * p.k_per_split;
      const int64_t kend = min(p.K, kb0 + p.k_per_split);
      for (int64_t k = kb0; k < kend; k += TC_BK) {
        if (lane == 0) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + Cfg::kABytes;
          mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          if (!A_MN) {
            tma_load_2d(&p.tmA, &full_bar[stage], sa, (int)k, mb * TC_BM);            // box {32 k, 128 m}
          } else if (p.a3d) {
            tma_load_3d(&p.tmA, &full_bar[stage], sa, 0, (int)k, mb * (TC_BM / 32));    // box {32 m, 32 k, 4 chunks}
          } else {
#pragma unroll
            for (int c = 0; c < TC_BM / 32; ++c)                                        // box {32 m, 32 k}
              tma_load_2d(&p.tmA, &full_bar[stage], sa + c * 4096, mb * TC_BM + c * 32, (int)k);
          }
          if (!B_MN) {
            tma_load_2d(&p.tmB, &full_bar[stage], sb, (int)k, nb * BN);               // box {32 k, BN n}
          } else if (p.b3d) {
            tma_load_3d(&p.tmB, &full_bar[stage], sb, 0, (int)k, nb * (BN / 32));       // box {32 n, 32 k, BN/32 chunks}
          } else {
#pragma unroll
            for (int c = 0; c < BN / 32; ++c)                                           // box {32 n, 32 k}
              tma_load_2d(&p.tmB, &full_bar[stage], sb + c * 4096, nb * BN + c * 32, (int)k);
          }
        }
        __syncwarp();
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // One lane runs the loop; its serial instruction stream shares an issue port with two epilogue warps, so the
    // per-k-block work is kept to waits + 4 MMAs + 1 commit: descriptors come from a precomputed base plus the
    // stage address, all tile bookkeeping is done once per tile.
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(TC_BM, BN, A_MN, B_MN);
      const uint64_t a_hi = !A_MN ? make_smem_desc(0, 16, 1024) : make_smem_desc(0, 4096, 512, 1);
      const uint64_t b_hi = !B_MN ? make_smem_desc(0, 16, 1024) : make_smem_desc(0, 4096, 512, 1);
      constexpr uint32_t a_step = !A_MN ? (32 >> 4) : (1024 >> 4);     // descriptor address units (16 B) per UMMA_K
      constexpr uint32_t b_step = !B_MN ? (32 >> 4) : (1024 >> 4);
      const uint32_t smem_s = smem_u32(smem);
      const uint32_t full_s = smem_u32(full_bar), empty_s = smem_u32(empty_bar);
      int stage = 0; uint32_t phase = 0;
      int64_t it = 0;
      for (int64_t t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
        int mb, nb, sp; decode(t, mb, nb, sp);
        const int64_t kb0 = (int64_t)sp * p.k_per_split;
        const int64_t kend = min(p.K, kb0 + p.k_per_split);
        const int nkb = (int)((kend - kb0 + TC_BK - 1) / TC_BK);
        const int as = (int)(it & 1); const uint32_t aphase = (uint32_t)((it >> 1) & 1);
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);
        uint32_t accumulate = 0;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait_addr(full_s + stage * 8u, phase);
          tc_fence_after();
          const uint32_t sa = smem_s + stage * Cfg::kStageBytes;
          const uint64_t adesc = a_hi | (uint64_t)((sa >> 4) & 0x3FFF);
          const uint64_t bdesc = b_hi | (uint64_t)(((sa + Cfg::kABytes) >> 4) & 0x3FFF);
          tc_mma_tf32(tmem_d, adesc, bdesc, idesc, accumulate);
          tc_mma_tf32(tmem_d, adesc + a_step, bdesc + b_step, idesc, 1);
          tc_mma_tf32(tmem_d, adesc + 2 * a_step, bdesc + 2 * b_step, idesc, 1);
          tc_mma_tf32(tmem_d, adesc + 3 * a_step, bdesc + 3 * b_step, idesc, 1);
          accumulate = 1;
          tc_commit_addr(empty_s + stage * 8u);          // smem stage reusable once these MMAs retire
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
        tc_commit(&tfull_bar[as]);                       // accumulator complete
      }
    }
  } else {
    // ===================== epilogue warps (8): quadrant = warp % 4, two warps per quadrant =====================
    const int quad = warp & 3;                       // TMEM lane quadrant this warp may access
    const int half = (warp - TC_EPI_WARP0) >> 2;     // which of the two warps of the quadrant
    float* stg = reinterpret_cast<float*>(smem + Cfg::kStages * Cfg::kStageBytes + 256) + (warp - TC_EPI_WARP0) * 32 * TC_STAGE_LD;
    float loss_acc = 0.f;
    int64_t ldaux; const float* auxp = epilogue_aux_ptr(p.ep, &ldaux);
    int64_t it = 0;
    for (int64_t t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      int mb, nb, sp; decode(t, mb, nb, sp);
      const int as = (int)(it & 1); const uint32_t aphase = (uint32_t)((it >> 1) & 1);
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const int64_t row0 = (int64_t)mb * TC_BM + quad * 32;
      float* cbase = p.C + (int64_t)sp * p.split_stride;
#pragma unroll 1
      for (int ch = half; ch < BN / 32; ch += 2) {
        uint32_t r[32];
        tc_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * BN + ch * 32), r);
#pragma unroll
        for (int j = 0; j < 32; ++j) stg[lane * TC_STAGE_LD + j] = __uint_as_float(r[j]);   // lane = row
        __syncwarp();
        const int64_t col = (int64_t)nb * BN + ch * 32 + lane;                               // lane = column
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
      if (lane == 0) mbar_arrive(&tempty_bar[as]);    // 8 arrivals (one per epilogue warp) free the accumulator
    }
    if (p.ep.loss_partials) {
      float w = warp_sum(loss_acc);
      if (lane == 0) epi_red[warp - TC_EPI_WARP0] = w;
      asm volatile("bar.sync 1, 256;" ::: "memory");  // epilogue warps only
      if (warp == TC_EPI_WARP0 && lane == 0) {
        float sacc = 0.f;
#pragma unroll
        for (int i = 0; i < TC_EPI_WARPS; ++i) sacc += epi_red[i];
        p.ep.loss_partials[blockIdx.x] = sacc;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(Cfg::kTmemCols) : "memory");
  }
}

template <int BN, bool A_MN, bool B_MN>
inline cudaError_t tc_launch_inst(const TcParams& p, int grid, cudaStream_t st) {
  static bool configured_dev[64] = {};        // the attribute is per device: one flag per device ordinal
  int dev_ = 0; cudaGetDevice(&dev_);
  bool& configured = configured_dev[dev_ & 63];
  auto kern = gemm_tc_kernel<BN, A_MN, B_MN>;
  if (!configured || dev_ >= 64) {      // (ordinals past the table are configured on every launch instead of aliasing)
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<BN>::kSmemBytes);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  kern<<<grid, TC_THREADS, TcCfg<BN>::kSmemBytes, st>>>(p);
  return cudaGetLastError();
}


template <int BN>
inline cudaError_t tc_launch_impl(bool a_mn, bool b_mn, const TcParams& p, int grid, cudaStream_t st) {
  if (!a_mn && !b_mn) return tc_launch_inst<BN, false, false>(p, grid, st);
  if (!a_mn && b_mn) return tc_launch_inst<BN, false, true>(p, grid, st);
  if (a_mn && !b_mn) return tc_launch_inst<BN, true, false>(p, grid, st);
  return tc_launch_inst<BN, true, true>(p, grid, st);
}

}  // namespace mmae
