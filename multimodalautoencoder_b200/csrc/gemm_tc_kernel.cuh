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
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* tm, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
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
  __shared__ float epi_red[4];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t tiles_mn = (int64_t)p.m_blocks * p.n_blocks;
  const int64_t num_tiles = tiles_mn * p.splits;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tmB) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], 4); }
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
          } else {
#pragma unroll
            for (int c = 0; c < TC_BM / 32; ++c)                                        // box {32 m, 32 k}
              tma_load_2d(&p.tmA, &full_bar[stage], sa + c * 4096, mb * TC_BM + c * 32, (int)k);
          }
          if (!B_MN) {
            tma_load_2d(&p.tmB, &full_bar[stage], sb, (int)k, nb * BN);               // box {32 k, BN n}
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
    constexpr uint32_t idesc = make_idesc_tf32(TC_BM, BN, A_MN, B_MN);
    int stage = 0; uint32_t phase = 0;
    int64_t it = 0;
    for (int64_t t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      int mb, nb, sp; decode(t, mb, nb, sp);
      const int64_t kb0 = (int64_t)sp * p.k_per_split;
      const int64_t kend = min(p.K, kb0 + p.k_per_split);
      const int as = (int)(it & 1); const uint32_t aphase = (uint32_t)((it >> 1) & 1);
      mbar_wait(&tempty_bar[as], aphase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);
      uint32_t accumulate = 0;
      for (int64_t k = kb0; k < kend; k += TC_BK) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint32_t sb = sa + Cfg::kABytes;
#pragma unroll
          for (int kk = 0; kk < TC_BK / TC_UMMA_K; ++kk) {
            // K-major: rows of 128 B, 8-row groups 1024 B apart; advance 32 B per UMMA_K inside the swizzle row.
            // MN-major (SW128, 32B atoms): atoms [4 k][32 mn] of 512 B; MN chunks 4096 B apart (LBO), k groups
            // 512 B apart (SBO); one UMMA_K = 8 spans two k groups = 1024 B.
            uint64_t adesc = !A_MN ? make_smem_desc(sa + kk * 32, 16, 1024) : make_smem_desc(sa + kk * 1024, 4096, 512, 1);
            uint64_t bdesc = !B_MN ? make_smem_desc(sb + kk * 32, 16, 1024) : make_smem_desc(sb + kk * 1024, 4096, 512, 1);
            tc_mma_tf32(tmem_d, adesc, bdesc, idesc, accumulate);
            accumulate = 1;
          }
          tc_commit(&empty_bar[stage]);              // smem stage reusable once these MMAs retire
        }
        __syncwarp();
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
      if (lane == 0) tc_commit(&tfull_bar[as]);       // accumulator complete
      __syncwarp();
    }
  } else {
    // ===================== epilogue warps =====================
    const int quad = warp & 3;                       // TMEM lane quadrant this warp may access
    float loss_acc = 0.f;
    int64_t it = 0;
    const bool vec_ok = ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0);
    for (int64_t t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      int mb, nb, sp; decode(t, mb, nb, sp);
      const int as = (int)(it & 1); const uint32_t aphase = (uint32_t)((it >> 1) & 1);
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const int64_t row = (int64_t)mb * TC_BM + quad * 32 + lane;
      float* crow = p.C + (int64_t)sp * p.split_stride + row * p.ldc;
#pragma unroll 1
      for (int ch = 0; ch < BN / 32; ++ch) {
        uint32_t r[32];
        tc_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * BN + ch * 32), r);
        const int64_t col0 = (int64_t)nb * BN + ch * 32;
        if (row < p.M && col0 < p.N) {
          if (vec_ok && col0 + 32 <= p.N) {
#pragma unroll
            for (int v = 0; v < 8; ++v) {
              float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
              if (p.ep.beta != 0.f) o = *reinterpret_cast<const float4*>(crow + col0 + v * 4);
              o.x = epilogue_apply(p.ep, row, col0 + v * 4 + 0, __uint_as_float(r[v * 4 + 0]), o.x, loss_acc);
              o.y = epilogue_apply(p.ep, row, col0 + v * 4 + 1, __uint_as_float(r[v * 4 + 1]), o.y, loss_acc);
              o.z = epilogue_apply(p.ep, row, col0 + v * 4 + 2, __uint_as_float(r[v * 4 + 2]), o.z, loss_acc);
              o.w = epilogue_apply(p.ep, row, col0 + v * 4 + 3, __uint_as_float(r[v * 4 + 3]), o.w, loss_acc);
              *reinterpret_cast<float4*>(crow + col0 + v * 4) = o;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (col0 + j < p.N) {
                float old = (p.ep.beta != 0.f) ? crow[col0 + j] : 0.f;
                crow[col0 + j] = epilogue_apply(p.ep, row, col0 + j, __uint_as_float(r[j]), old, loss_acc);
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);    // 4 arrivals (one per epilogue warp) free the accumulator
    }
    if (p.ep.loss_partials) {
      float w = warp_sum(loss_acc);
      if (lane == 0) epi_red[quad] = w;
      asm volatile("bar.sync 1, 128;" ::: "memory");  // epilogue warps only
      if (warp == TC_EPI_WARP0 && lane == 0)
        p.ep.loss_partials[blockIdx.x] = epi_red[0] + epi_red[1] + epi_red[2] + epi_red[3];
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
  static bool configured = false;
  auto kern = gemm_tc_kernel<BN, A_MN, B_MN>;
  if (!configured) {
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
