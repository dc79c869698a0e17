// Device side of the grouped weight-gradient GEMM (design notes in wgrad_group.cuh).
#include "wgrad_group.cuh"
#include "gemm_tc_kernel.cuh"

namespace mmae {

using WgCfg = TcCfg<WG_BN>;

__global__ void __launch_bounds__(TC_THREADS, 1) wgrad_group_kernel(const __grid_constant__ WgParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + WgCfg::kStages * WgCfg::kStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + WgCfg::kStages;
  uint64_t* tfull_bar = bars + 2 * WgCfg::kStages;
  uint64_t* tempty_bar = bars + 2 * WgCfg::kStages + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * WgCfg::kStages + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t num_items = (int64_t)p.tiles * p.splits;

  if (warp == 0 && lane == 0)
    for (int q = 0; q < p.nprob; ++q) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&p.pr[q].tmA) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&p.pr[q].tmB) : "memory");
    }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < WgCfg::kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], TC_EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(WgCfg::kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // item -> (problem, m block, n block, batch slice): slices outermost, so the CTAs resident together work on the same
  // batch rows and the tiles of one problem share their operand in L2
  auto decode = [&](int64_t t, int& q, int& mb, int& nb, int& sp) {
    sp = (int)(t / p.tiles);
    const int r = (int)(t - (int64_t)sp * p.tiles);
    q = 0;
    while (q + 1 < p.nprob && r >= p.pr[q + 1].tile0) ++q;
    const int local = r - p.pr[q].tile0;
    mb = local / p.pr[q].n_blocks;
    nb = local - mb * p.pr[q].n_blocks;
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    int stage = 0; uint32_t phase = 0;
    for (int64_t t = blockIdx.x; t < num_items; t += gridDim.x) {
      int q, mb, nb, sp; decode(t, q, mb, nb, sp);
      const WgProblem& P = p.pr[q];
      const int64_t kb0 = (int64_t)sp * p.k_per_split;
      const int64_t kend = min(p.K, kb0 + p.k_per_split);
      for (int64_t k = kb0; k < kend; k += TC_BK) {
        if (lane == 0) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * WgCfg::kStageBytes;
          uint8_t* sb = sa + WgCfg::kABytes;
          mbar_expect_tx(&full_bar[stage], WgCfg::kStageBytes);
          if (P.a3d) {
            tma_load_3d(&P.tmA, &full_bar[stage], sa, 0, (int)k, mb * (TC_BM / 32));      // box {32 m, 32 k, 4 chunks}
          } else {
#pragma unroll
            for (int c = 0; c < TC_BM / 32; ++c)                                          // box {32 m, 32 k}
              tma_load_2d(&P.tmA, &full_bar[stage], sa + c * 4096, mb * TC_BM + c * 32, (int)k);
          }
          if (P.b3d) {
            tma_load_3d(&P.tmB, &full_bar[stage], sb, 0, (int)k, nb * (WG_BN / 32));
          } else {
#pragma unroll
            for (int c = 0; c < WG_BN / 32; ++c)
              tma_load_2d(&P.tmB, &full_bar[stage], sb + c * 4096, nb * WG_BN + c * 32, (int)k);
          }
        }
        __syncwarp();
        if (++stage == WgCfg::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(TC_BM, WG_BN, true, true);
      const uint64_t d_hi = make_smem_desc(0, 4096, 512, 1);          // MN-major, 128B swizzle with 32-byte atoms
      constexpr uint32_t step = 1024 >> 4;
      const uint32_t smem_s = smem_u32(smem);
      const uint32_t full_s = smem_u32(full_bar), empty_s = smem_u32(empty_bar);
      int stage = 0; uint32_t phase = 0;
      int64_t it = 0;
      for (int64_t t = blockIdx.x; t < num_items; t += gridDim.x, ++it) {
        int q, mb, nb, sp; decode(t, q, mb, nb, sp);
        const int64_t kb0 = (int64_t)sp * p.k_per_split;
        const int64_t kend = min(p.K, kb0 + p.k_per_split);
        const int nkb = (int)((kend - kb0 + TC_BK - 1) / TC_BK);
        const int as = (int)(it & 1); const uint32_t aphase = (uint32_t)((it >> 1) & 1);
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * WG_BN);
        uint32_t accumulate = 0;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait_addr(full_s + stage * 8u, phase);
          tc_fence_after();
          const uint32_t sa = smem_s + stage * WgCfg::kStageBytes;
          const uint64_t adesc = d_hi | (uint64_t)((sa >> 4) & 0x3FFF);
          const uint64_t bdesc = d_hi | (uint64_t)(((sa + WgCfg::kABytes) >> 4) & 0x3FFF);
          tc_mma_tf32(tmem_d, adesc, bdesc, idesc, accumulate);
          tc_mma_tf32(tmem_d, adesc + step, bdesc + step, idesc, 1);
          tc_mma_tf32(tmem_d, adesc + 2 * step, bdesc + 2 * step, idesc, 1);
          tc_mma_tf32(tmem_d, adesc + 3 * step, bdesc + 3 * step, idesc, 1);
          accumulate = 1;
          tc_commit_addr(empty_s + stage * 8u);
          if (++stage == WgCfg::kStages) { stage = 0; phase ^= 1; }
        }
        tc_commit(&tfull_bar[as]);
      }
    }
  } else {
    // ===================== epilogue warps: partial tile -> its split-K slice =====================
    const int quad = warp & 3;
    const int half = (warp - TC_EPI_WARP0) >> 2;
    float* stg = reinterpret_cast<float*>(smem + WgCfg::kStages * WgCfg::kStageBytes + 256) + (warp - TC_EPI_WARP0) * 32 * TC_STAGE_LD;
    Epilogue ep; ep.mode = EPI_PLAIN; ep.beta = 0.f; ep.keep = 1.f;
    float dummy = 0.f;
    int64_t it = 0;
    for (int64_t t = blockIdx.x; t < num_items; t += gridDim.x, ++it) {
      int q, mb, nb, sp; decode(t, q, mb, nb, sp);
      const WgProblem& P = p.pr[q];
      const int as = (int)(it & 1); const uint32_t aphase = (uint32_t)((it >> 1) & 1);
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const int64_t row0 = (int64_t)mb * TC_BM + quad * 32;
      float* cbase = P.ws + (int64_t)sp * P.M * P.N;
#pragma unroll 1
      for (int ch = half; ch < WG_BN / 32; ch += 2) {
        if (nb * WG_BN + ch * 32 >= P.N || row0 >= P.M) continue;       // warp-uniform: nothing of this chunk is inside the matrix
        uint32_t r[32];
        tc_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * WG_BN + ch * 32), r);
#pragma unroll
        for (int j = 0; j < 32; ++j) stg[lane * TC_STAGE_LD + j] = __uint_as_float(r[j]);
        __syncwarp();
        const int64_t col = (int64_t)nb * WG_BN + ch * 32 + lane;
        if (col < P.N) {
          EpiRowCtx c;
          c.cp = cbase + row0 * P.N + col; c.ap = nullptr; c.ldc = P.N; c.ldaux = 0; c.stg = stg + lane;
          c.nrows = (int)min((int64_t)32, (int64_t)P.M - row0);
          c.bias_v = 0.f; c.beta = 0.f; c.grow0 = row0; c.col = col; c.colsum_out = nullptr;
          epi_rows<EPI_PLAIN, 0, false>(ep, c, dummy);
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(WgCfg::kTmemCols) : "memory");
  }
}

cudaError_t wgrad_group_launch(const WgParams& p, int grid, cudaStream_t st) {
  static bool configured_dev[64] = {};
  int dev_ = 0; cudaGetDevice(&dev_);
  bool& configured = configured_dev[dev_ & 63];
  if (!configured || dev_ >= 64) {      // (ordinals past the table are configured on every launch instead of aliasing)
    cudaError_t e = cudaFuncSetAttribute(wgrad_group_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WgCfg::kSmemBytes);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  wgrad_group_kernel<<<grid, TC_THREADS, WgCfg::kSmemBytes, st>>>(p);
  return cudaGetLastError();
}

}  // namespace mmae
