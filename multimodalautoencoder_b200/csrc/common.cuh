// Shared device helpers: Philox4x32-10, activations, the fused GEMM epilogue.
// sm_100a only.  Reference lines cited are in /root/reference/multimodal_autoencoder.py.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mmae_b200.h"

namespace mmae {

// ------------------------------------------------------------------ Philox4x32-10
// Host twin: oracle/philox_host.py (must match bit for bit).
constexpr uint32_t kStreamBatch = 1, kStreamZero = 2, kStreamMod = 3, kStreamEps = 4, kStreamDrop = 16;

struct Philox4 { uint32_t x, y, z, w; };

__host__ __device__ __forceinline__ Philox4 philox4x32(uint64_t index, uint32_t stream, uint32_t step,
                                                       uint64_t seed) {
  uint32_t c0 = (uint32_t)index, c1 = (uint32_t)(index >> 32), c2 = stream, c3 = step;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return Philox4{c0, c1, c2, c3};
}

__device__ __forceinline__ uint32_t philox_word(uint64_t element, uint32_t stream, uint32_t step,
                                                uint64_t seed) {
  Philox4 p = philox4x32(element >> 2, stream, step, seed);
  uint32_t l = (uint32_t)element & 3u;
  return l == 0 ? p.x : (l == 1 ? p.y : (l == 2 ? p.z : p.w));
}

__host__ __device__ __forceinline__ uint32_t mulhi_u32(uint32_t a, uint32_t n) {
  return (uint32_t)(((uint64_t)a * n) >> 32);
}

// ------------------------------------------------------------------ activations (:477-497)
__device__ __forceinline__ float act_fwd(int act, float z) {
  switch (act) {
    case MMAE_ACT_RELU: return fmaxf(z, 0.f);
    case MMAE_ACT_TANH: return tanhf(z);
    case MMAE_ACT_SOFTSIGN: return z / (1.f + fabsf(z));
    case MMAE_ACT_SOFTPLUS: return fmaxf(z, 0.f) + log1pf(expf(-fabsf(z)));
    default: return z;
  }
}

__device__ __forceinline__ float act_fwd_fast(int act, float z) {
  switch (act) {
    case MMAE_ACT_RELU: return fmaxf(z, 0.f);
    case MMAE_ACT_TANH: { float e = __expf(-2.f * fabsf(z)); float t = __fdividef(1.f - e, 1.f + e); return z >= 0.f ? t : -t; }
    case MMAE_ACT_SOFTSIGN: return __fdividef(z, 1.f + fabsf(z));
    case MMAE_ACT_SOFTPLUS: return fmaxf(z, 0.f) + __logf(1.f + __expf(-fabsf(z)));
    default: return z;
  }
}

// act'(z) expressed through the stored output a = act(z)
__device__ __forceinline__ float act_bwd_from_output(int act, float a) {
  switch (act) {
    case MMAE_ACT_RELU: return a > 0.f ? 1.f : 0.f;
    case MMAE_ACT_TANH: return 1.f - a * a;
    case MMAE_ACT_SOFTSIGN: { float t = 1.f - fabsf(a); return t * t; }   // 1/(1+|z|)^2
    case MMAE_ACT_SOFTPLUS: return 1.f - expf(-a);                         // sigmoid(z)
    default: return 1.f;
  }
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// ------------------------------------------------------------------ noise descriptor (:649-702)
struct NoiseView {
  const uint32_t* zero_bits;   // [B, zw]
  const uint32_t* mod_bits;    // [B]
  const uint8_t* col_mod;      // [F] column -> modality
  int zw;                      // ceil(F/32)
  float mask_with;
  int enabled;
};

__device__ __forceinline__ float noisy_value(const NoiseView& nv, int64_t row, int col, float x) {
  if (!nv.enabled) return x;
  uint32_t zb = __ldg(nv.zero_bits + row * nv.zw + (col >> 5));
  uint32_t mb = __ldg(nv.mod_bits + row);
  if ((mb >> __ldg(nv.col_mod + col)) & 1u) return nv.mask_with;   // block mask wins (:695 after :683)
  if ((zb >> (col & 31)) & 1u) return 0.f;
  return x;
}

// ------------------------------------------------------------------ fused epilogue
enum EpiMode : int {
  EPI_PLAIN = 0,       // C = acc + beta*C                                  (wgrad, raw dgrad)
  EPI_BIAS_ACT = 1,    // C = drop(act(acc + bias))                         (:467-474, :510-517)
  EPI_LOSS_TRAIN = 2,  // l = acc+bias; loss += f(l,T); C = dLoss/dl        (:381-390 + backward seed)
  EPI_LOSS_PRED = 3,   // l = acc+bias; loss += f(l,T); C = decoded_X       (:378/:390)
  EPI_DGRAD = 4        // C = (acc + beta*C) * act'(S) * dropmask/keep      (appendix B)
};

struct Epilogue {
  int mode;
  const float* bias;      // [N] or null
  int act;
  float beta;
  // dropout (forward: applied to output; dgrad: applied with saved activation)
  float keep;             // 1.0 = off
  uint32_t keep_thr;      // ceil(keep * 2^24)
  uint32_t drop_stream;   // kStreamDrop + slot
  const uint32_t* step;   // device-resident Philox step (StepState)
  uint64_t seed;
  int64_t drop_width;     // logical width of the dropped activation (element = row*width+col)
  int64_t row0;           // global index of local row 0 (data-parallel shards share one stream)
  // loss modes
  const float* target;    // [M, ldt]
  int64_t ldt;
  const int64_t* aux_rows; // optional (row-layout tcgen05 epilogue only): row r of the target is target[aux_rows[r]] -- the loss
                          // reads its clean rows straight from a resident dataset through the sampled index list
  int loss;               // mmae_loss
  float* loss_partials;   // one float per CTA (deterministic two-stage reduction), may be null
  float* colsum_partials; // tcgen05 family only: [ceil(M/32), N] column sums of the stored values per 32-row group
  // dgrad mode
  const float* saved;     // stored activations h = drop(act(z)), [M, lds]
  int64_t lds;
  // fill-in (whole-network kernel only, EPI_LOSS_PRED): C = block missing in the input row ? decoded_X : target
  // (data_funcs.py:310-381); fill_bits[row] has bit m set when modality m of the row is missing
  const uint32_t* fill_bits;
  const uint8_t* fill_col_mod;   // [N] column -> modality
};

// Which auxiliary matrix the epilogue reads at (row, col): the loss target or the saved activation.
__device__ __forceinline__ const float* epilogue_aux_ptr(const Epilogue& ep, int64_t* ld) {
  if (ep.mode == EPI_LOSS_TRAIN || ep.mode == EPI_LOSS_PRED) { *ld = ep.ldt; return ep.target; }
  if (ep.mode == EPI_DGRAD) { *ld = ep.lds; return ep.saved; }
  *ld = 0; return nullptr;
}

// Applies the epilogue to one accumulator; returns the value to store.  The caller supplies the bias
// value of the column, the auxiliary value (target / saved activation) and the current C value (only
// meaningful when beta != 0), so that it can hoist and batch those loads; `loss_acc` collects the
// per-thread loss contribution.  FAST selects hardware-approximate exp/log/divide (tcgen05 family,
// tf32 tolerance); the CUDA-core fp32 family keeps the accurate forms.
template <bool FAST = false>
__device__ __forceinline__ float epilogue_apply(const Epilogue& ep, int64_t row, int64_t col, float acc,
                                                float bias_v, float aux, float c_old, float& loss_acc) {
  switch (ep.mode) {
    case EPI_PLAIN:
      return ep.beta != 0.f ? acc + ep.beta * c_old : acc;
    case EPI_BIAS_ACT: {
      float v = acc + bias_v;
      v = FAST ? act_fwd_fast(ep.act, v) : act_fwd(ep.act, v);
      if (ep.keep < 1.f) {
        uint32_t w = philox_word((uint64_t)(row + ep.row0) * (uint64_t)ep.drop_width + (uint64_t)col, ep.drop_stream, __ldg(ep.step), ep.seed);
        v = ((w >> 8) < ep.keep_thr) ? v / ep.keep : 0.f;
      }
      return v;
    }
    case EPI_LOSS_TRAIN:
    case EPI_LOSS_PRED: {
      const float l = acc + bias_v;
      const float x = aux;
      const bool has_t = ep.target != nullptr;
      float out;
      if (ep.loss == MMAE_LOSS_SIGMOID_CE) {
        float s;
        if (FAST) {       // one ex2 + one lg2 + one rcp: e = exp(-|l|); sigmoid = (l >= 0 ? 1 : e) / (1 + e)
          float e = __expf(-fabsf(l));
          float inv = __fdividef(1.f, 1.f + e);
          s = l >= 0.f ? inv : e * inv;
          if (has_t) loss_acc += fmaxf(l, 0.f) - l * x + __logf(1.f + e);
        } else {
          s = sigmoidf_(l);
          if (has_t) loss_acc += fmaxf(l, 0.f) - l * x + log1pf(expf(-fabsf(l)));
        }
        out = (ep.mode == EPI_LOSS_TRAIN) ? (s - x) : s;
      } else if (ep.loss == MMAE_LOSS_RMSE) {
        float d = l - x;
        if (has_t) loss_acc += d * d;
        out = (ep.mode == EPI_LOSS_TRAIN) ? d : l;      // unscaled; 1/(N*rmse) is applied in Adam
      } else {
        if (has_t) loss_acc += -x * logf(l);
        out = (ep.mode == EPI_LOSS_TRAIN) ? (-x / l) : l;
      }
      return out;
    }
    case EPI_DGRAD: {
      float g = ep.beta != 0.f ? acc + ep.beta * c_old : acc;
      float h = aux;
      if (ep.keep < 1.f) {
        uint32_t w = philox_word((uint64_t)(row + ep.row0) * (uint64_t)ep.drop_width + (uint64_t)col, ep.drop_stream, __ldg(ep.step), ep.seed);
        if ((w >> 8) < ep.keep_thr) { g = g / ep.keep; h = h * ep.keep; } else { return 0.f; }
      }
      return g * act_bwd_from_output(ep.act, h);
    }
  }
  return acc;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace mmae
