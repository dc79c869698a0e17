// Streaming / reduction kernels around the GEMMs.  All are HBM-bound: 128-bit accesses where the
// layout allows, grids sized as a multiple of the SM count, fixed-order (deterministic) reductions.
// Reference lines cited are in /root/reference/multimodal_autoencoder.py unless a file is named.
#pragma once
#include "common.cuh"

namespace mmae {

// ------------------------------------------------------------------ noise descriptor (:668-702)
// One thread per row draws the row's modality mask; one warp-strided loop draws the n_zero
// column indices (WITH replacement, :682) and ORs them into the row's bitmap.
struct NoiseGenArgs {
  uint32_t* zero_bits; uint32_t* mod_bits;
  int64_t batch, row0;
  int num_feats, zw, n_zero, num_mod;
  int mode, num_types, num_drop;
  uint32_t thresholds[8]; uint32_t type_masks[8];
  const uint32_t* step; uint64_t seed;     // step lives in device memory (StepState) so that captured graphs replay
};

__global__ void noise_gen_kernel(const NoiseGenArgs a) {
  // one warp per row
  int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (row >= a.batch) return;
  int64_t grow = row + a.row0;
  uint32_t* zb = a.zero_bits + row * a.zw;
  for (int w = lane; w < a.zw; w += 32) zb[w] = 0u;
  __syncwarp();
  int q = (a.n_zero + 3) >> 2;                      // Philox counters per row
  for (int c = lane; c < q; c += 32) {
    Philox4 p = philox4x32((uint64_t)grow * q + c, kStreamZero, __ldg(a.step), a.seed);
    uint32_t wv[4] = {p.x, p.y, p.z, p.w};
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      int j = c * 4 + l;
      if (j < a.n_zero) {
        uint32_t col = mulhi_u32(wv[l], (uint32_t)a.num_feats);
        atomicOr(zb + (col >> 5), 1u << (col & 31));
      }
    }
  }
  if (lane == 0) {
    Philox4 p = philox4x32((uint64_t)grow, kStreamMod, __ldg(a.step), a.seed);
    uint32_t mb = 0u;
    if (a.mode == MMAE_NOISE_INTELLIGENT) {         // categorical over noise types (:689-695)
      int k = 0;
      for (int t = 0; t < a.num_types - 1; ++t) k += (p.x >= a.thresholds[t]) ? 1 : 0;
      mb = a.type_masks[k];
    } else {                                        // randint(0, M) num_drop times (:698-700)
      uint32_t wv[4] = {p.x, p.y, p.z, p.w};
      for (int d = 0; d < a.num_drop; ++d) mb |= 1u << mulhi_u32(wv[d], (uint32_t)a.num_mod);
    }
    a.mod_bits[row] = mb;
  }
}

// noisy_X materialised (add_noise_to_batch's return value).  One warp per row: the row's bitmap words and modality
// mask are read once, X streams through as float4 (F % 4 == 0) with 128-bit coalesced accesses.
__global__ void noise_apply_kernel(const float* __restrict__ X, float* __restrict__ out, int64_t batch,
                                   int num_feats, NoiseView nv) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < batch; r += nwarps) {
    const float* x = X + r * (int64_t)num_feats;
    float* o = out + r * (int64_t)num_feats;
    const uint32_t mb = nv.enabled ? __ldg(nv.mod_bits + r) : 0u;
    const uint32_t* zb = nv.zero_bits + r * nv.zw;
    if ((num_feats & 3) == 0) {
      for (int c4 = lane; c4 < (num_feats >> 2); c4 += 32) {
        float4 v = __ldg(reinterpret_cast<const float4*>(x) + c4);
        const int c = c4 << 2;
        if (nv.enabled) {
          const uint32_t z = __ldg(zb + (c >> 5)) >> (c & 31);
          const uchar4 m = *reinterpret_cast<const uchar4*>(nv.col_mod + c);
          v.x = ((mb >> m.x) & 1u) ? nv.mask_with : ((z & 1u) ? 0.f : v.x);
          v.y = ((mb >> m.y) & 1u) ? nv.mask_with : ((z & 2u) ? 0.f : v.y);
          v.z = ((mb >> m.z) & 1u) ? nv.mask_with : ((z & 4u) ? 0.f : v.z);
          v.w = ((mb >> m.w) & 1u) ? nv.mask_with : ((z & 8u) ? 0.f : v.w);
        }
        reinterpret_cast<float4*>(o)[c4] = v;
      }
    } else {
      for (int c = lane; c < num_feats; c += 32) o[c] = noisy_value(nv, r, c, __ldg(x + c));
    }
  }
}

// One pass that produces everything a train step needs from its batch: optional row sampling (Philox indices into a
// resident dataset, data_funcs.py:167) + the row's noise descriptor (:668-702) + the clean batch (loss target) + the
// noisy batch (first GEMM operand and its weight gradient).  One warp per row; the row's zero bitmap is built in shared
// memory, so the descriptor is drawn and applied without a round trip through global memory.  Bit-identical to
// philox_indices_kernel + gather_rows_kernel + noise_gen_kernel + noise_apply_kernel run one after the other.
struct SampleNoiseArgs {
  NoiseGenArgs g;                 // descriptor parameters; g.zero_bits / g.mod_bits still receive the descriptor
  const float* src;               // dataset [n_rows, F] (gather) or the batch itself [batch, F]
  uint32_t n_rows;                // > 0: sample row indices into src;  0: src is the batch
  const int64_t* idx_in;          // optional host-supplied indices (n_rows > 0)
  const int64_t* view;            // optional: row list of the current training view (a cross-validation fold): sampled
                                  // index j means dataset row view[j];  n_rows is then the length of the view
  int64_t* idx_out;               // optional: the dataset rows that were gathered
  float* clean_out;               // gathered clean batch (null when src is the batch)
  float* noisy_out;               // noisy batch
  const uint8_t* col_mod; float mask_with;
};
constexpr int SN_WARPS = 8;
constexpr int SN_MAX_ZW = 128;      // F <= 4096
__global__ void __launch_bounds__(SN_WARPS * 32) sample_noise_kernel(const SampleNoiseArgs a) {
  __shared__ uint32_t zsm[SN_WARPS][SN_MAX_ZW];
  const int wl = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t nwarps = (int64_t)gridDim.x * SN_WARPS;
  const uint32_t step = __ldg(a.g.step);
  const int F = a.g.num_feats, zw = a.g.zw;
  for (int64_t row = (int64_t)blockIdx.x * SN_WARPS + wl; row < a.g.batch; row += nwarps) {
    const int64_t grow = row + a.g.row0;
    uint32_t* zb = zsm[wl];
    for (int w = lane; w < zw; w += 32) zb[w] = 0u;
    __syncwarp();
    const int q = (a.g.n_zero + 3) >> 2;
    for (int c = lane; c < q; c += 32) {
      Philox4 p = philox4x32((uint64_t)grow * q + c, kStreamZero, step, a.g.seed);
      uint32_t wv[4] = {p.x, p.y, p.z, p.w};
#pragma unroll
      for (int l = 0; l < 4; ++l) {
        int j = c * 4 + l;
        if (j < a.g.n_zero) {
          uint32_t col = mulhi_u32(wv[l], (uint32_t)F);
          atomicOr(zb + (col >> 5), 1u << (col & 31));
        }
      }
    }
    uint32_t mb = 0u;
    {
      Philox4 p = philox4x32((uint64_t)grow, kStreamMod, step, a.g.seed);      // every lane draws the same word: no shuffle needed
      if (a.g.mode == MMAE_NOISE_INTELLIGENT) {
        int k = 0;
        for (int t = 0; t < a.g.num_types - 1; ++t) k += (p.x >= a.g.thresholds[t]) ? 1 : 0;
        mb = a.g.type_masks[k];
      } else {
        uint32_t wv[4] = {p.x, p.y, p.z, p.w};
        for (int d = 0; d < a.g.num_drop; ++d) mb |= 1u << mulhi_u32(wv[d], (uint32_t)a.g.num_mod);
      }
    }
    int64_t srow = row;
    if (a.n_rows) {
      srow = a.idx_in ? a.idx_in[row] : (int64_t)mulhi_u32(philox_word((uint64_t)grow, kStreamBatch, step, a.g.seed), a.n_rows);
      if (a.view) srow = __ldg(a.view + srow);
      if (lane == 0 && a.idx_out) a.idx_out[row] = srow;
    }
    __syncwarp();
    if (lane == 0) a.g.mod_bits[row] = mb;
    for (int w = lane; w < zw; w += 32) a.g.zero_bits[row * zw + w] = zb[w];
    const float* x = a.src + srow * (int64_t)F;
    float* o = a.noisy_out + row * (int64_t)F;
    float* cl = a.clean_out ? a.clean_out + row * (int64_t)F : nullptr;
    if ((F & 3) == 0) {
      for (int c4 = lane; c4 < (F >> 2); c4 += 32) {
        float4 v = __ldg(reinterpret_cast<const float4*>(x) + c4);
        if (cl) reinterpret_cast<float4*>(cl)[c4] = v;
        const int c = c4 << 2;
        const uint32_t z = zb[c >> 5] >> (c & 31);
        const uchar4 m = *reinterpret_cast<const uchar4*>(a.col_mod + c);
        v.x = ((mb >> m.x) & 1u) ? a.mask_with : ((z & 1u) ? 0.f : v.x);
        v.y = ((mb >> m.y) & 1u) ? a.mask_with : ((z & 2u) ? 0.f : v.y);
        v.z = ((mb >> m.z) & 1u) ? a.mask_with : ((z & 4u) ? 0.f : v.z);
        v.w = ((mb >> m.w) & 1u) ? a.mask_with : ((z & 8u) ? 0.f : v.w);
        reinterpret_cast<float4*>(o)[c4] = v;
      }
    } else {
      for (int c = lane; c < F; c += 32) {
        float v = __ldg(x + c);
        if (cl) cl[c] = v;
        const bool zz = (zb[c >> 5] >> (c & 31)) & 1u;
        o[c] = ((mb >> a.col_mod[c]) & 1u) ? a.mask_with : (zz ? 0.f : v);
      }
    }
    __syncwarp();
  }
}

// out[c][r] = in[r][c]  (K-major shadows of the weights for the tcgen05 forward GEMMs)
__global__ void transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int rows, int cols) {
  __shared__ float t[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    int r = r0 + i, c = c0 + threadIdx.x;
    t[i][threadIdx.x] = (r < rows && c < cols) ? __ldg(in + (int64_t)r * cols + c) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    int c = c0 + i, r = r0 + threadIdx.x;
    if (r < rows && c < cols) out[(int64_t)c * rows + r] = t[threadIdx.x][i];
  }
}

// All stale weight shadows in ONE launch (a step used to issue one small transpose per weight matrix).
struct TransposeGroup {
  static constexpr int kMax = 16;
  int n;
  int64_t off[kMax];        // offset of the variable in P / PT
  int rows[kMax], cols[kMax];
  int tile0[kMax + 1];      // first 32 x 32 tile of each variable in the flattened grid
};
__global__ void transpose_group_kernel(const float* __restrict__ P, float* __restrict__ PT, const TransposeGroup g) {
  __shared__ float t[32][33];
  int v = 0;
  while (v + 1 < g.n && (int)blockIdx.x >= g.tile0[v + 1]) ++v;
  const int rows = g.rows[v], cols = g.cols[v];
  const int tiles_c = (cols + 31) / 32;
  const int tl = blockIdx.x - g.tile0[v];
  const int c0 = (tl % tiles_c) * 32, r0 = (tl / tiles_c) * 32;
  const float* in = P + g.off[v]; float* out = PT + g.off[v];
  for (int i = threadIdx.y; i < 32; i += 8) {
    int r = r0 + i, c = c0 + threadIdx.x;
    t[i][threadIdx.x] = (r < rows && c < cols) ? __ldg(in + (int64_t)r * cols + c) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    int c = c0 + i, r = r0 + threadIdx.x;
    if (r < rows && c < cols) out[(int64_t)c * rows + r] = t[threadIdx.x][i];
  }
}

// ------------------------------------------------------------------ batch gather (data_funcs.py:167-168)
__global__ void gather_rows_kernel(const float* __restrict__ src, const int64_t* __restrict__ idx,
                                   float* __restrict__ dst, int64_t batch, int width) {
  // one warp per row, float4 when possible
  int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (row >= batch) return;
  const float* s = src + idx[row] * (int64_t)width;
  float* d = dst + row * (int64_t)width;
  if ((width & 3) == 0) {
    const float4* s4 = reinterpret_cast<const float4*>(s); float4* d4 = reinterpret_cast<float4*>(d);
    for (int i = lane; i < (width >> 2); i += 32) d4[i] = __ldg(s4 + i);
  } else {
    for (int i = lane; i < width; i += 32) d[i] = __ldg(s + i);
  }
}

// idx[i] = dataset row of batch row i: Philox draw in [0, n_rows) (or the given index), mapped through the optional view
__global__ void philox_indices_kernel(int64_t* idx, int64_t batch, int64_t first, uint32_t n_rows,
                                      const uint32_t* step, uint64_t seed, const int64_t* given, const int64_t* view) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= batch) return;
  int64_t j = given ? given[i] : (int64_t)mulhi_u32(philox_word((uint64_t)(i + first), kStreamBatch, __ldg(step), seed), n_rows);
  idx[i] = view ? __ldg(view + j) : j;
}

// ------------------------------------------------------------------ column sums (bias gradients)
// stage 1: grid (ceil(N/32), S); block 32x8; partial[s][n] = sum over this split's rows.
__global__ void colsum_partial_kernel(const float* __restrict__ D, int64_t rows, int n, int64_t ld,
                                      float* __restrict__ partial, int splits) {
  __shared__ float sm[8][33];
  int col = blockIdx.x * 32 + threadIdx.x;
  int64_t per = (rows + splits - 1) / splits;
  int64_t r0 = (int64_t)blockIdx.y * per, r1 = min(rows, r0 + per);
  float s = 0.f;
  if (col < n)
    for (int64_t r = r0 + threadIdx.y; r < r1; r += 8) s += __ldg(D + r * ld + col);
  sm[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && col < n) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += sm[i][threadIdx.x];
    partial[(int64_t)blockIdx.y * n + col] = t;
  }
}
__global__ void colsum_final_kernel(const float* __restrict__ partial, int n, int splits, float* __restrict__ out) {
  int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= n) return;
  float t = 0.f;
  for (int s = 0; s < splits; ++s) t += partial[(int64_t)s * n + col];
  out[col] = t;
}

// split-K partial slices -> C (+ beta*C), fixed order
__global__ void splitk_reduce_kernel(const float* __restrict__ ws, int64_t mn, int splits, float* __restrict__ C,
                                     int64_t n, int64_t ldc, float beta) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < mn; i += (int64_t)gridDim.x * blockDim.x) {
    float t = 0.f;
    for (int s = 0; s < splits; ++s) t += __ldg(ws + (int64_t)s * mn + i);
    int64_t r = i / n, c = i - r * n;
    float* p = C + r * ldc + c;
    *p = beta != 0.f ? t + beta * (*p) : t;
  }
}

// ------------------------------------------------------------------ loss finalisation
// Sums per-CTA partials in a fixed order (double) and writes scalars; single block.
// sums[0] = recon loss sum / sumsq, sums[1] = KL sum, ...; see engine.cu for slot use.
__global__ void reduce_partials_kernel(const float* __restrict__ partials, int64_t n, double* __restrict__ out_slot,
                                       int accumulate) {
  __shared__ double sm[256];
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += (double)partials[i];
  sm[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out_slot = accumulate ? (*out_slot + sm[0]) : sm[0];
}

// ------------------------------------------------------------------ VAE (:372-375, :402-406)
struct VaeArgs {
  const float* mu; const float* lv; float* eps; float* emb; float* kl_partials;
  int64_t batch, row0; int E; const uint32_t* step; uint64_t seed; int gen_eps;
};
__global__ void vae_sample_kernel(const VaeArgs a) {
  __shared__ float red[8];
  int64_t total = a.batch * (int64_t)a.E;
  float kl = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    float e;
    if (a.gen_eps) {
      Philox4 p = philox4x32((uint64_t)(i + a.row0 * a.E), kStreamEps, __ldg(a.step), a.seed);
      float u1 = ((float)(p.x >> 8) + 1.0f) * (1.0f / 16777216.0f);
      float u2 = (float)(p.y >> 8) * (1.0f / 16777216.0f);
      e = sqrtf(-2.f * logf(u1)) * cospif(2.f * u2);
      a.eps[i] = e;
    } else {
      e = a.eps[i];
    }
    float lv = a.lv[i];
    float z = a.mu[i] + e * expf(lv);
    a.emb[i] = z;
    kl += -0.5f * (1.f + 2.f * lv - z * z - expf(2.f * lv));
  }
  float w = warp_sum(kl);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = w;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += red[i];
    a.kl_partials[blockIdx.x] = s;
  }
}
// g_mu = g_e + kl_w*emb ; g_lv = g_mu*eps*exp(lv) + kl_w*(exp(2 lv) - 1)      (appendix B)
// kl_w = 1/B_global on the reconstruction step, 0 on the classification step.
__global__ void vae_grad_kernel(float* __restrict__ g_e /*in: dL/demb, out: g_mu*/, float* __restrict__ g_lv,
                                const float* __restrict__ emb, const float* __restrict__ lv,
                                const float* __restrict__ eps, int64_t total, float kl_w) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    float l = lv[i];
    float g = g_e[i] + kl_w * emb[i];
    g_e[i] = g;
    g_lv[i] = g * eps[i] * expf(l) + kl_w * (expf(2.f * l) - 1.f);
  }
}

// ------------------------------------------------------------------ head loss (:431-452)
struct HeadLossArgs {
  const float* logits; const float* labels; float* delta; float* probs; int32_t* preds;
  float* partials;      // [gridDim.x * 2]: loss sum, correct count
  int64_t batch; int C; int loss; float inv_count;   // 1/(B_global*C) or 1/B_global
};
__global__ void head_loss_kernel(const HeadLossArgs a) {
  __shared__ float red[2][8];
  float ls = 0.f, correct = 0.f;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < a.batch; r += (int64_t)gridDim.x * blockDim.x) {
    const float* l = a.logits + r * a.C;
    if (a.loss == MMAE_HEAD_SIGMOID_CE) {
      for (int c = 0; c < a.C; ++c) {
        float v = l[c], s = sigmoidf_(v);
        int p = v > 0.f ? 1 : 0;                       // round-half-even(sigmoid) (:448)
        if (a.probs) a.probs[r * a.C + c] = s;
        if (a.preds) a.preds[r * a.C + c] = p;
        if (a.labels) {
          float y = a.labels[r * a.C + c];
          ls += fmaxf(v, 0.f) - v * y + log1pf(expf(-fabsf(v)));
          correct += (p == (int)y) ? 1.f : 0.f;        // tf.cast(float->int32) truncates (:425)
          if (a.delta) a.delta[r * a.C + c] = (s - y) * a.inv_count;
        }
      }
    } else {
      float mx = l[0]; int am = 0;
      for (int c = 1; c < a.C; ++c) if (l[c] > mx) { mx = l[c]; am = c; }
      float se = 0.f;
      for (int c = 0; c < a.C; ++c) se += expf(l[c] - mx);
      if (a.probs) for (int c = 0; c < a.C; ++c) a.probs[r * a.C + c] = sigmoidf_(l[c]);
      if (a.preds) a.preds[r] = am;                    // argmax (:450)
      if (a.labels) {
        int y = min(max((int)a.labels[r], 0), a.C - 1);   // the Python layer rejects out-of-range labels; never index past the row
        ls += mx + logf(se) - l[y];
        correct += (am == y) ? 1.f : 0.f;
        if (a.delta) for (int c = 0; c < a.C; ++c)
          a.delta[r * a.C + c] = (expf(l[c] - mx) / se - (c == y ? 1.f : 0.f)) * a.inv_count;
      }
    }
  }
  float w0 = warp_sum(ls), w1 = warp_sum(correct);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = w0; red[1][threadIdx.x >> 5] = w1; }
  __syncthreads();
  if (threadIdx.x == 0 && a.partials) {
    float s0 = 0.f, s1 = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { s0 += red[0][i]; s1 += red[1][i]; }
    a.partials[blockIdx.x] = s0;
    a.partials[gridDim.x + blockIdx.x] = s1;
  }
}

// ------------------------------------------------------------------ fused scale + L2 + TF-Adam (:411, :443)
// tf.train.AdamOptimizer._apply_dense:  a_t = lr*sqrt(1-b2^t)/(1-b1^t);  m += (g-m)(1-b1);
// v += (g^2-v)(1-b2);  theta -= a_t * m / (sqrt(v)+eps).   g = scale*G + l2[seg]*theta.
struct AdamSeg { int64_t begin; float l2; float pad; int rows, cols; };      // rows / cols of a 2-D variable (cols = 0: vector)
struct AdamArgs {
  float* P; const float* G; float* M; float* V;
  int64_t begin, end;            // flat range inside P / G;  M, V are indexed from 0 at `begin`
  const AdamSeg* segs; int nsegs;
  const double* sums;            // device scalars (loss sums after the optional allreduce)
  int scale_mode;                // 0: 1.0   1: RMSE 1/sqrt(N*sumsq), N = n_elems   2: clip_by_global_norm (sums[6] = ||g||^2)
  float clip_norm;
  double n_elems;
  const float* alpha;            // lr * sqrt(1 - b2^t) / (1 - b1^t), written by adam_prep_kernel for this step
  float b1, b2, eps;
  double* scalars_out;           // MMAE_S_* slots, written by thread 0
  float* PT;                     // when set: the K-major shadow of every 2-D variable is refreshed in the same pass
};
__global__ void adam_kernel(const AdamArgs a) {
  float scale = 1.f;
  if (a.scale_mode == 1) scale = (float)(1.0 / sqrt(a.n_elems * a.sums[0]));
  float clip = 1.f;              // tf.clip_by_global_norm: g * clip_norm / max(||g||, clip_norm), g INCLUDING the L2 term
  if (a.scale_mode == 2) { const double nrm = sqrt(a.sums[6]); clip = (float)((double)a.clip_norm / fmax(nrm, (double)a.clip_norm)); }
  if (blockIdx.x == 0 && threadIdx.x == 0 && a.scalars_out) a.scalars_out[MMAE_S_GRAD_SCALE] = scale;
  const float alpha = __ldg(a.alpha);
  int64_t n = a.end - a.begin;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t gi = a.begin + i;
    int lo = 0, hi = a.nsegs - 1;                 // last segment with begin <= gi
    while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (a.segs[mid].begin <= gi) lo = mid; else hi = mid - 1; }
    float p = a.P[gi];
    float g = (scale * a.G[gi] + a.segs[lo].l2 * p) * clip;
    float m = a.M[i], v = a.V[i];
    m += (g - m) * (1.f - a.b1);
    v += (g * g - v) * (1.f - a.b2);
    a.M[i] = m; a.V[i] = v;
    const float pn = p - alpha * m / (sqrtf(v) + a.eps);
    a.P[gi] = pn;
    if (a.PT && a.segs[lo].cols > 0) {            // small models: scattered 4-byte stores are cheaper than a launch
      const int64_t loc = gi - a.segs[lo].begin;
      const int cols = a.segs[lo].cols, rows = a.segs[lo].rows;
      if (loc < (int64_t)rows * cols) {
        const int r = (int)(loc / cols), c = (int)(loc - (int64_t)r * cols);
        a.PT[a.segs[lo].begin + (int64_t)c * rows + r] = pn;
      }
    }
  }
}

// ||G + l2 * P||^2 over [begin, end): per-block partials (fixed order), summed by reduce_partials_kernel into sums[6]
__global__ void grad_sqnorm_kernel(const float* __restrict__ P, const float* __restrict__ G, int64_t begin, int64_t end,
                                   const AdamSeg* __restrict__ segs, int nsegs, float* __restrict__ partials) {
  __shared__ float red[8];
  float acc = 0.f;
  for (int64_t gi = begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; gi < end; gi += (int64_t)gridDim.x * blockDim.x) {
    int lo = 0, hi = nsegs - 1;
    while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (segs[mid].begin <= gi) lo = mid; else hi = mid - 1; }
    const float g = G[gi] + segs[lo].l2 * P[gi];
    acc += g * g;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
    partials[blockIdx.x] = t;
  }
}

// Large models: Adam over 32 x 32 tiles of every variable of the range, so that the K-major shadow of a weight matrix
// (what the tcgen05 forward GEMMs read) is rewritten through shared memory in the same pass -- coalesced both ways -- instead
// of by transpose launches at the top of the next forward.  Vectors ride along as [1, n] matrices (no shadow).
struct AdamTiles {
  static constexpr int kMax = 24;
  int n;
  int64_t off[kMax]; int rows[kMax], cols[kMax]; float l2[kMax]; int shadow[kMax];
  int tile0[kMax + 1];
};
__global__ void __launch_bounds__(256) adam_tiled_kernel(const AdamArgs a, const AdamTiles g) {
  __shared__ float t[32][33];
  float scale = 1.f;
  if (a.scale_mode == 1) scale = (float)(1.0 / sqrt(a.n_elems * a.sums[0]));
  float clip = 1.f;
  if (a.scale_mode == 2) { const double nrm = sqrt(a.sums[6]); clip = (float)((double)a.clip_norm / fmax(nrm, (double)a.clip_norm)); }
  if (blockIdx.x == 0 && threadIdx.x == 0 && threadIdx.y == 0 && a.scalars_out) a.scalars_out[MMAE_S_GRAD_SCALE] = scale;
  const float alpha = __ldg(a.alpha);
  int v = 0;
  while (v + 1 < g.n && (int)blockIdx.x >= g.tile0[v + 1]) ++v;
  const int rows = g.rows[v], cols = g.cols[v];
  const int tiles_c = (cols + 31) / 32;
  const int tl = blockIdx.x - g.tile0[v];
  const int c0 = (tl % tiles_c) * 32, r0 = (tl / tiles_c) * 32;
  const int64_t off = g.off[v];
  const float l2 = g.l2[v];
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    float pn = 0.f;
    if (r < rows && c < cols) {
      const int64_t gi = off + (int64_t)r * cols + c;
      const int64_t k = gi - a.begin;
      const float p = a.P[gi];
      const float gr = (scale * a.G[gi] + l2 * p) * clip;
      float m = a.M[k], vv = a.V[k];
      m += (gr - m) * (1.f - a.b1);
      vv += (gr * gr - vv) * (1.f - a.b2);
      a.M[k] = m; a.V[k] = vv;
      pn = p - alpha * m / (sqrtf(vv) + a.eps);
      a.P[gi] = pn;
    }
    t[i][threadIdx.x] = pn;
  }
  if (!g.shadow[v]) return;
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < rows && c < cols) a.PT[off + (int64_t)c * rows + r] = t[threadIdx.x][i];
  }
}

// Per-step state kept in DEVICE memory so that a captured CUDA graph of the train step replays unchanged: the Philox
// step index every random draw is keyed by, and the two optimizers' step counts with their bias-corrected rates.
struct StepState { uint32_t step; uint32_t pad; long long t[2]; float alpha[2]; };
__global__ void advance_step_kernel(StepState* s) { s->step += 1u; }
__global__ void adam_prep_kernel(StepState* s, int opt, double lr, double b1, double b2) {
  const long long t = ++s->t[opt];                      // TF ApplyAdam: t starts at 1 (:411, :443)
  s->alpha[opt] = (float)(lr * sqrt(1.0 - pow(b2, (double)t)) / (1.0 - pow(b1, (double)t)));
}

// ------------------------------------------------------------------ gradient assembly (small-config train step)
// One launch between the grouped weight-gradient GEMM and Adam:  block 0 sums the loss partials in a fixed order and
// does the per-step scalar bookkeeping (finalize_scalars_kernel's job);  weight segments sum their split-K slices in
// slice order;  bias segments sum the per-32-row column-sum partials the chain kernels left behind (8 row lanes per
// column, combined in a fixed tree).  Everything is fixed-order, hence deterministic.
struct GaSeg {
  int64_t g_off, count;      // destination range in G
  const float* src;          // weights: slices [nslices][count];  biases: partials [nslices][count]
  int nslices;
  int kind;                  // 0 weight, 1 bias
  int block0, nblocks;
};
constexpr int GA_MAX_SEGS = 24;
struct GaArgs {
  GaSeg seg[GA_MAX_SEGS]; int nseg;
  float* G;
  const float* loss_partials; int n_loss_partials; double* loss_sum_out;     // optional (null / 0: sums already there)
  int do_finalize;
};
// ------------------------------------------------------------------ fill-in (data_funcs.py:310-381)
// miss[r] bit m set iff sum(x[r, s_m:e_m]) == -(e_m - s_m); one warp per row, lanes stride the block.
__global__ void missing_bits_kernel(const float* __restrict__ X, int64_t batch, int num_feats,
                                    const int32_t* __restrict__ starts, int num_mod, uint32_t* __restrict__ miss) {
  // one warp per row; 256 columns per pass, all 8 loads of a lane in flight before any reduction.  Modalities are
  // contiguous column ranges in increasing order, so one running carry covers a block that spans passes.
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  // persistent: a fixed grid of warps strides over the rows (10 M one-row blocks would be bound by block launch rate)
  for (int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < batch; row += nwarps) {
    const float* x = X + row * (int64_t)num_feats;
    uint32_t bits = 0u;
    int m = 0;
    float carry = 0.f;
    for (int p0 = 0; p0 < num_feats && m < num_mod; p0 += 256) {
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) { const int c = p0 + e * 32 + lane; v[e] = c < num_feats ? __ldg(x + c) : 0.f; }
      const int pend = min(num_feats, p0 + 256);
      while (m < num_mod && starts[m] < pend) {
        const int s = starts[m], e1 = starts[m + 1];
        float part = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) { const int c = p0 + e * 32 + lane; part += (c >= s && c < e1) ? v[e] : 0.f; }
        part = warp_sum(part);
        if (e1 <= pend) {
          if (carry + part == -(float)(e1 - s)) bits |= 1u << m;
          carry = 0.f; ++m;
        } else { carry += part; break; }
      }
    }
    if (lane == 0) miss[row] = bits;
  }
}
__global__ void fill_select_kernel(const float* __restrict__ X, const float* __restrict__ recon,
                                   const uint32_t* __restrict__ miss, const uint8_t* __restrict__ col_mod,
                                   float* __restrict__ out, int64_t batch, int num_feats) {
  int64_t total = batch * (int64_t)num_feats;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / num_feats; int c = (int)(i - r * num_feats);
    out[i] = ((__ldg(miss + r) >> __ldg(col_mod + c)) & 1u) ? __ldg(recon + i) : __ldg(X + i);
  }
}

// ------------------------------------------------------------------ per-modality reconstruction error (:1189-1216)
// get_reconstruction_loss_per_modality as ONE batched pass: the M masked copies of the rows are stacked into one
// [M * n, F] batch (copy m has modality m set to the literal -1.0, :1203), one forward reconstructs all of them, and
// the squared error of every copy's own block is reduced in-kernel.
__global__ void modality_mask_batch_kernel(const float* __restrict__ X, float* __restrict__ out, int64_t n, int num_feats,
                                           int num_mod, const uint8_t* __restrict__ col_mod) {
  const int64_t total = (int64_t)num_mod * n * num_feats;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % num_feats);
    const int64_t rr = i / num_feats;
    const int m = (int)(rr / n);
    const int64_t r = rr - (int64_t)m * n;
    out[i] = (__ldg(col_mod + c) == m) ? -1.0f : __ldg(X + r * num_feats + c);
  }
}
// partial[m * gridDim.x + blockIdx.x] = sum over this block's rows of copy m, columns of block m, of (X - recon)^2
__global__ void modality_sse_kernel(const float* __restrict__ X, const float* __restrict__ recon, int64_t n, int num_feats,
                                    const int32_t* __restrict__ starts, int num_mod, double* __restrict__ partial) {
  __shared__ double red[8];
  const int m = blockIdx.y;
  const int s = starts[m], e = starts[m + 1], w = e - s;
  double acc = 0.0;
  const int64_t total = n * (int64_t)w;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / w; const int c = s + (int)(i - r * w);
    const float d = __ldg(X + r * num_feats + c) - __ldg(recon + ((int64_t)m * n + r) * num_feats + c);
    acc += (double)d * (double)d;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
    partial[(int64_t)m * gridDim.x + blockIdx.x] = t;
  }
}
// sse[m] += sum of the block partials (fixed order); with finish != 0: rmse[m] = sqrt(sse[m] / (rows * width_m))
__global__ void modality_sse_reduce_kernel(const double* __restrict__ partial, int nblocks, double* __restrict__ sse,
                                           const int32_t* __restrict__ starts, int64_t rows_total, int finish, int reset) {
  const int m = blockIdx.x;
  if (threadIdx.x != 0) return;
  double t = reset ? 0.0 : sse[m];
  for (int i = 0; i < nblocks; ++i) t += partial[(int64_t)m * nblocks + i];
  const int w = starts[m + 1] - starts[m];
  sse[m] = finish ? (w > 0 && rows_total > 0 ? sqrt(t / ((double)rows_total * w)) : nan("")) : t;
}

// ------------------------------------------------------------------ small glue kernels
__global__ void elementwise_epilogue_kernel(float* d, int64_t rows, int cols, Epilogue ep) {
  int64_t total = rows * (int64_t)cols;
  float dummy = 0.f;
  int64_t ldaux; const float* auxp = epilogue_aux_ptr(ep, &ldaux);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / cols, c = i - r * cols;
    float aux = auxp ? auxp[r * ldaux + c] : 0.f;
    d[i] = epilogue_apply<false>(ep, r, c, d[i], 0.f, aux, 0.f, dummy);
  }
}
__global__ void pack_sums_kernel(double* sums, float* tail, int to_tail) {
  int i = threadIdx.x;
  if (i < 8) { if (to_tail) tail[i] = (float)sums[i]; else sums[i] = (double)tail[i]; }
}
struct StepState;
struct FinalizeArgs {
  const double* sums; double* scalars; int loss; int variational; double n_elems, batch, head_count; int do_recon, do_head;
  // end-of-step bookkeeping folded into this single-thread kernel (small-batch steps are launch-latency bound):
  StepState* state; int prep_opt; double lr, b1, b2; int advance;      // prep_opt >= 0: ++t, alpha for that optimizer; advance: ++step
};
__device__ __forceinline__ void finalize_scalars_body(const FinalizeArgs& a);
__global__ void finalize_scalars_kernel(FinalizeArgs a) { finalize_scalars_body(a); }

__device__ __forceinline__ void finalize_scalars_body(const FinalizeArgs& a) {
  if (a.do_recon) {
    if (a.loss == MMAE_LOSS_RMSE) { a.scalars[MMAE_S_SUMSQ] = a.sums[0]; a.scalars[MMAE_S_RECON_LOSS] = sqrt(a.sums[0] / a.n_elems); }
    else { a.scalars[MMAE_S_SUMSQ] = 0.0; a.scalars[MMAE_S_RECON_LOSS] = a.sums[0]; }
    a.scalars[MMAE_S_KL_MEAN] = a.variational ? a.sums[1] / a.batch : 0.0;
  }
  if (a.do_head) {
    a.scalars[MMAE_S_HEAD_LOSS] = a.sums[2] / a.head_count;
    a.scalars[MMAE_S_HEAD_ACC] = a.sums[3] / a.head_count;
  }
  if (a.state) {
    if (a.prep_opt >= 0) {
      const long long t = ++a.state->t[a.prep_opt];
      a.state->alpha[a.prep_opt] = (float)(a.lr * sqrt(1.0 - pow(a.b2, (double)t)) / (1.0 - pow(a.b1, (double)t)));
    }
    if (a.advance) a.state->step += 1u;
  }
}

constexpr int GA_THREADS = 1024;
__global__ void __launch_bounds__(GA_THREADS) grad_assemble_kernel(const GaArgs a, const FinalizeArgs fin) {
  if (blockIdx.x == 0) {
    __shared__ double sm[GA_THREADS];
    if (a.loss_partials) {
      double s = 0.0;
      for (int i = threadIdx.x; i < a.n_loss_partials; i += GA_THREADS) s += (double)a.loss_partials[i];
      sm[threadIdx.x] = s;
      __syncthreads();
      for (int o = GA_THREADS / 2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
        __syncthreads();
      }
      if (threadIdx.x == 0) *a.loss_sum_out = sm[0];
    }
    if (threadIdx.x == 0 && a.do_finalize) finalize_scalars_body(fin);
    return;
  }
  int si = 0;
  while (si + 1 < a.nseg && (int)blockIdx.x >= a.seg[si + 1].block0) ++si;
  const GaSeg& sg = a.seg[si];
  const int lb = blockIdx.x - sg.block0;
  if (sg.kind == 0) {
    for (int64_t i = (int64_t)lb * GA_THREADS + threadIdx.x; i < sg.count; i += (int64_t)sg.nblocks * GA_THREADS) {
      float t = 0.f;
#pragma unroll 6
      for (int s = 0; s < sg.nslices; ++s) t += __ldg(sg.src + (int64_t)s * sg.count + i);
      a.G[sg.g_off + i] = t;
    }
  } else {
    // 32 columns x 32 row lanes; each lane walks its groups 8 loads at a time, then a fixed-order tree over the lanes
    __shared__ float red[32][33];
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int64_t col = (int64_t)lb * 32 + cx;
    float t = 0.f;
    if (col < sg.count) {
      int g = ry;
      for (; g + 7 * 32 < sg.nslices; g += 8 * 32) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldg(sg.src + (int64_t)(g + u * 32) * sg.count + col);
#pragma unroll
        for (int u = 0; u < 8; ++u) t += v[u];
      }
      for (; g < sg.nslices; g += 32) t += __ldg(sg.src + (int64_t)g * sg.count + col);
    }
    red[ry][cx] = t;
    __syncthreads();
    if (ry == 0 && col < sg.count) {
      float u = 0.f;
#pragma unroll
      for (int i = 0; i < 32; ++i) u += red[i][cx];
      a.G[sg.g_off + col] = u;
    }
  }
}

}  // namespace mmae
