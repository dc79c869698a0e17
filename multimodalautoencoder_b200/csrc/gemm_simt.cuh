// CUDA-core fp32 GEMM with the fused MMAE prologue (block-mask noise on the X operand) and
// epilogue (bias / activation / dropout / loss / act').  This is the exact-fp32 family:
// it serves MMAE_PREC_FP32 engines and every shape the tcgen05 family cannot take
// (leading dimensions not a multiple of 4 floats, the 2-/3-wide head, batch < 128).
//
//   C[M,N] = opA(A) . opB(B),  row-major C;  TA: A stored [K,M];  TB: B stored [N,K].
//
// 64x64x16 block tile, 256 threads, 4x4 register tile per thread; every global access is
// bounds-checked so any M, N, K works (the reference's placeholders are shapeless,
// multimodal_autoencoder.py:351-352).
#pragma once
#include "common.cuh"

namespace mmae {

constexpr int SG_BM = 64, SG_BN = 64, SG_BK = 32, SG_THREADS = 256;

struct GemmArgs {
  int64_t M, N, K;
  const float* A; int64_t lda;
  const float* B; int64_t ldb;
  float* C; int64_t ldc;
  NoiseView noise;      // applied to A's stored (row, col) = (batch row, feature) when enabled
  int noise_aligned32;       // modality boundaries are multiples of 32 columns (two-SM tcgen05 patch fast path)
  Epilogue ep;
  int splits;           // split-K over blockIdx.z: the CTAs of one output tile write their partial tiles to `ws`, the last one
  int64_t k_per_split;  //   to arrive (per-tile counter) adds them up in slice order and runs the epilogue: one launch,
  float* ws;            //   deterministic, and small-batch GEMMs (M = 20..100: one or two tiles) still fill the SMs
  int* tile_counters;   //   zero on entry, reset by the reducing CTA
};

template <bool TA, bool TB>
__global__ void __launch_bounds__(SG_THREADS) gemm_simt_kernel(const GemmArgs g) {
  __shared__ float As[SG_BK][SG_BM + 4];
  __shared__ float Bs[SG_BK][SG_BN + 4];
  __shared__ float red[SG_THREADS / 32];

  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * SG_BM;     // M tiles on grid.x (no 65535 limit)
  const int64_t n0 = (int64_t)blockIdx.y * SG_BN;
  const int tx = tid & 15, ty = tid >> 4;     // 16 x 16 threads; thread owns rows ty*4.., cols tx*4..

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int64_t kbeg = g.splits > 1 ? (int64_t)blockIdx.z * g.k_per_split : 0;
  const int64_t kend = g.splits > 1 ? min(g.K, kbeg + g.k_per_split) : g.K;
  // Register-prefetched K loop: the global loads of tile k+1 are in flight while tile k is multiplied out of shared
  // memory.  Small-batch steps (M = 20..100) run a handful of CTAs per GEMM, so each CTA's loop latency is the step time.
  constexpr int LPT = SG_BM * SG_BK / SG_THREADS;      // elements of each operand tile per thread (8)
  float ra[LPT], rb[LPT];
  auto fetch = [&](int64_t k0) {
#pragma unroll
    for (int i = 0; i < LPT; ++i) {
      int m, k;
      if (!TA) { k = tid & 31; m = (tid >> 5) + 8 * i; }      // K contiguous in memory
      else     { m = tid & 63; k = (tid >> 6) + 4 * i; }      // M contiguous in memory
      int64_t gm = m0 + m, gk = k0 + k;
      float v = 0.f;
      if (gm < g.M && gk < kend) {
        if (!TA) { v = __ldg(g.A + gm * g.lda + gk); v = noisy_value(g.noise, gm, (int)gk, v); }
        else     { v = __ldg(g.A + gk * g.lda + gm); v = noisy_value(g.noise, gk, (int)gm, v); }
      }
      ra[i] = v;
    }
#pragma unroll
    for (int i = 0; i < LPT; ++i) {
      int n, k;
      if (!TB) { n = tid & 63; k = (tid >> 6) + 4 * i; }      // N contiguous
      else     { k = tid & 31; n = (tid >> 5) + 8 * i; }      // K contiguous
      int64_t gn = n0 + n, gk = k0 + k;
      float v = 0.f;
      if (gn < g.N && gk < kend) v = !TB ? __ldg(g.B + gk * g.ldb + gn) : __ldg(g.B + gn * g.ldb + gk);
      rb[i] = v;
    }
  };
  auto stash = [&]() {
#pragma unroll
    for (int i = 0; i < LPT; ++i) {
      if (!TA) As[tid & 31][(tid >> 5) + 8 * i] = ra[i]; else As[(tid >> 6) + 4 * i][tid & 63] = ra[i];
      if (!TB) Bs[(tid >> 6) + 4 * i][tid & 63] = rb[i]; else Bs[tid & 31][(tid >> 5) + 8 * i] = rb[i];
    }
  };
  if (kbeg < kend) fetch(kbeg);
  for (int64_t k0 = kbeg; k0 < kend; k0 += SG_BK) {
    stash();
    __syncthreads();
    if (k0 + SG_BK < kend) fetch(k0 + SG_BK);
#pragma unroll
    for (int k = 0; k < SG_BK; ++k) {
      float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

  if (g.splits > 1) {
    // partial tile -> workspace [tile][slice][64 x 64]; the last CTA of the tile sums the slices in order
    __shared__ int is_last;
    const int64_t tile = (int64_t)blockIdx.y * gridDim.x + blockIdx.x;
    float* wt = g.ws + (tile * g.splits + blockIdx.z) * (SG_BM * SG_BN);
#pragma unroll
    for (int i = 0; i < 4; ++i)
      *reinterpret_cast<float4*>(wt + (ty * 4 + i) * SG_BN + tx * 4) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      const int prev = atomicAdd(g.tile_counters + tile, 1);
      is_last = prev == g.splits - 1;
      if (is_last) g.tile_counters[tile] = 0;            // ready for the next launch
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    const float* w0 = g.ws + tile * g.splits * (SG_BM * SG_BN);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int z = 0; z < g.splits; ++z) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 v = __ldcg(reinterpret_cast<const float4*>(w0 + (int64_t)z * (SG_BM * SG_BN) + (ty * 4 + i) * SG_BN + tx * 4));
        acc[i][0] += v.x; acc[i][1] += v.y; acc[i][2] += v.z; acc[i][3] += v.w;
      }
    }
  }
  float* Cout = g.C;
  float loss_acc = 0.f;
  int64_t ldaux; const float* auxp = epilogue_aux_ptr(g.ep, &ldaux);
  float bias_v[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int64_t c = n0 + tx * 4 + j;
    bias_v[j] = (g.ep.bias && c < g.N) ? __ldg(g.ep.bias + c) : 0.f;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t r = m0 + ty * 4 + i;
    if (r >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int64_t c = n0 + tx * 4 + j;
      if (c >= g.N) continue;
      float* p = Cout + r * g.ldc + c;
      float old = (g.ep.beta != 0.f) ? *p : 0.f;
      float aux = auxp ? __ldg(auxp + r * ldaux + c) : 0.f;
      *p = epilogue_apply<false>(g.ep, r, c, acc[i][j], bias_v[j], aux, old, loss_acc);
    }
  }
  if (g.ep.loss_partials) {             // fixed-order block reduction -> one partial per CTA
    float w = warp_sum(loss_acc);
    if ((tid & 31) == 0) red[tid >> 5] = w;
    __syncthreads();
    if (tid == 0) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < SG_THREADS / 32; ++i) s += red[i];
      g.ep.loss_partials[(int64_t)blockIdx.y * gridDim.x + blockIdx.x] = s;
    }
  }
}

inline int64_t gemm_simt_num_ctas(int64_t M, int64_t N) {
  return ((M + SG_BM - 1) / SG_BM) * ((N + SG_BN - 1) / SG_BN);
}

// Split count for a CUDA-core wgrad whose tile grid would leave most SMs idle (K = batch is the long dimension).
inline int simt_pick_splits(int64_t M, int64_t N, int64_t K, int num_sms) {
  const int64_t tiles = gemm_simt_num_ctas(M, N);
  if (tiles >= num_sms || K < 128) return 1;
  int64_t s = (num_sms + tiles - 1) / tiles;
  const int64_t max_s = K / 64;         // >= 2 k-iterations per slice
  if (s > max_s) s = max_s;
  if (s > 64) s = 64;
  return (int)(s < 1 ? 1 : s);
}

inline cudaError_t launch_gemm_simt(bool ta, bool tb, const GemmArgs& g, cudaStream_t st) {
  dim3 grid((unsigned)((g.M + SG_BM - 1) / SG_BM), (unsigned)((g.N + SG_BN - 1) / SG_BN), (unsigned)(g.splits > 1 ? g.splits : 1));
  if (grid.y > 65535u) return cudaErrorInvalidConfiguration;
  if (!ta && !tb) gemm_simt_kernel<false, false><<<grid, SG_THREADS, 0, st>>>(g);
  else if (!ta && tb) gemm_simt_kernel<false, true><<<grid, SG_THREADS, 0, st>>>(g);
  else if (ta && !tb) gemm_simt_kernel<true, false><<<grid, SG_THREADS, 0, st>>>(g);
  else gemm_simt_kernel<true, true><<<grid, SG_THREADS, 0, st>>>(g);
  return cudaGetLastError();
}

}  // namespace mmae
