// Device side of the whole-network chain kernel (design notes in chain_tc.cuh).
#include "chain_tc.cuh"
#include "gemm_tc_kernel.cuh"
#include <stdio.h>

namespace mmae {

// ------------------------------------------------------------------ PTX wrappers specific to the chain
// 20 warps share 4 issue ports and most of them are waiting at any time: a bare try_wait loop re-issues every ~35
// cycles (measured: 3/4 of all executed instructions were spin iterations), so waits carry a suspend-time hint and
// the hardware parks the warp until the phase flips.
__device__ __forceinline__ void mbar_wait_parked(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u) : "memory");
  } while (!done);
}
__device__ __forceinline__ void mbar_wait_s(uint32_t bar_saddr, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar_saddr), "r"(parity), "r"(0x989680u) : "memory");
  } while (!done);
}
__device__ __forceinline__ void tc_commit_s(uint32_t bar_saddr) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_saddr) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
        "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
        "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t to_tf32(float v) {
  uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v)); return r;
}
__host__ __device__ inline uint32_t idesc_tf32_rt(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// L2 policies: rows of X are read twice per tile (operand of the first layer, then loss target / fill-in source) and
// the weights by every tile -> evict_last; the second read of X and all outputs are streaming -> evict_first.
__device__ __forceinline__ uint64_t l2_policy_evict_last() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ uint64_t l2_policy_evict_first() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ void tma_load_2d_hint(const CUtensorMap* tm, uint64_t* bar, void* dst, int c0, int c1, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(pol) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(tm), "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_2d_hint(const CUtensorMap* tm, const void* src, int c0, int c1, uint64_t pol) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;"
               ::"l"(tm), "r"(smem_u32(src)), "r"(c0), "r"(c1), "l"(pol) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// One epilogue warp's staging: two [32 rows][32 cols] fp32 tiles in the TMA 128B-swizzle layout.  Thread = row:
// 16-byte chunk q of row r lives at r*128 + ((q ^ (r & 7)) << 4)  (conflict-free for the 8 lanes of a quarter warp).
struct EpiStage {
  uint8_t* buf[2];
  uint64_t* aux_bar[2];
  uint32_t aux_phase[2];
  uint32_t uses;          // tiles handed to TMA so far (alternates the two buffers)
};
// .ftz forms: without them every ex2 / lg2 / rcp drags a denormal-rescaling FSETP + FMUL + FSEL sequence along
__device__ __forceinline__ float ex2_ftz(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_ftz(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int ACT> __device__ __forceinline__ float act_chain_t(float z) {
  if (ACT == MMAE_ACT_SOFTSIGN) return z * rcp_ftz(1.f + fabsf(z));
  return act_fast_t<ACT>(z);
}
// log(1 + e) for e in [0, 1]: degree-7 minimax polynomial (max abs error 3e-7), FMA pipe instead of a third MUFU
__device__ __forceinline__ float log1p_unit(float e) {
  float r = 0.010243828408420086f;
  r = fmaf(r, e, -0.053267478942871094f);
  r = fmaf(r, e, 0.13198965787887573f);
  r = fmaf(r, e, -0.22396689653396606f);
  r = fmaf(r, e, 0.327511727809906f);
  r = fmaf(r, e, -0.4993339478969574f);
  r = fmaf(r, e, 0.9999702572822571f);
  return fmaf(r, e, 2.2159764512252877e-07f);
}
// Explicit shared-space accesses: through generic pointers the compiler emits LD.E / ST.E and, unable to prove that
// the bias table and the staging tile do not alias, serialises every load behind the previous store.
__device__ __forceinline__ uint32_t row_chunk(uint32_t tile_saddr, int lane, int q) {
  return tile_saddr + lane * 128 + ((q ^ (lane & 7)) << 4);
}
__device__ __forceinline__ float4 lds128(uint32_t a) {            // ordered with the barriers / stores around it
  float4 v; asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a)); return v;
}
__device__ __forceinline__ float4 lds128_ro(uint32_t a) {         // read-only table (bias): free to be hoisted
  float4 v; asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a)); return v;
}
__device__ __forceinline__ void sts128(uint32_t a, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float lds32(uint32_t a) {
  float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v;
}

// ------------------------------------------------------------------ epilogue of a non-final op, one 32-column chunk
// v = drop(act(acc + bias)); TMEM <- tf32(v) in place; optional global copy through a swizzled tile + TMA store.
template <int ACT, bool DROP>
__device__ __forceinline__ void chain_act_chunk(const ChainOp& o, const CUtensorMap* tmO, uint32_t taddr, int col0, int64_t tile_row0,
                                                int quad, int lane, uint32_t bias_s, EpiStage& es) {
  uint32_t r[32];
  tc_ld32(taddr, r);
  const uint32_t tile = smem_u32(es.buf[0]);
  if (o.has_out) {                       // the previous store has finished reading the (single) staging tile
    if (lane == 0) bulk_wait_read<0>();
    __syncwarp();
  }
  const int64_t grow = tile_row0 + quad * 32 + lane + o.ep.row0;      // this thread's global row (dropout stream)
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 b = lds128_ro(bias_s + (col0 + q * 4) * 4);     // warp-uniform address: broadcast
    float v[4] = {__uint_as_float(r[q * 4 + 0]) + b.x, __uint_as_float(r[q * 4 + 1]) + b.y,
                  __uint_as_float(r[q * 4 + 2]) + b.z, __uint_as_float(r[q * 4 + 3]) + b.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      v[e] = act_chain_t<ACT>(v[e]);
      if (DROP) {
        uint32_t w = philox_word((uint64_t)grow * (uint64_t)o.ep.drop_width + (uint64_t)(col0 + q * 4 + e), o.ep.drop_stream, __ldg(o.ep.step), o.ep.seed);
        v[e] = ((w >> 8) < o.ep.keep_thr) ? v[e] / o.ep.keep : 0.f;
      }
      r[q * 4 + e] = to_tf32(v[e]);
    }
    if (o.has_out) sts128(row_chunk(tile, lane, q), make_float4(v[0], v[1], v[2], v[3]));
  }
  tc_st32(taddr, r);
  if (o.has_out) {
    fence_async_smem();
    __syncwarp();
    if (lane == 0) { tma_store_2d(tmO, es.buf[0], col0, (int)(tile_row0 + quad * 32)); bulk_commit(); }
    es.uses++;
  }
}

template <bool DROP>
__device__ __forceinline__ void chain_act_dispatch(const ChainOp& o, const CUtensorMap* tmO, uint32_t taddr, int col0, int64_t tile_row0,
                                                   int quad, int lane, uint32_t bias_s, EpiStage& es) {
  switch (o.ep.act) {
    case MMAE_ACT_RELU: chain_act_chunk<MMAE_ACT_RELU, DROP>(o, tmO, taddr, col0, tile_row0, quad, lane, bias_s, es); break;
    case MMAE_ACT_TANH: chain_act_chunk<MMAE_ACT_TANH, DROP>(o, tmO, taddr, col0, tile_row0, quad, lane, bias_s, es); break;
    case MMAE_ACT_SOFTSIGN: chain_act_chunk<MMAE_ACT_SOFTSIGN, DROP>(o, tmO, taddr, col0, tile_row0, quad, lane, bias_s, es); break;
    case MMAE_ACT_SOFTPLUS: chain_act_chunk<MMAE_ACT_SOFTPLUS, DROP>(o, tmO, taddr, col0, tile_row0, quad, lane, bias_s, es); break;
    default: chain_act_chunk<MMAE_ACT_LINEAR, DROP>(o, tmO, taddr, col0, tile_row0, quad, lane, bias_s, es); break;
  }
}

// ------------------------------------------------------------------ backward chain: epilogue of a non-final dgrad op
// v = (delta . W^T) * act'(h) * dropmask/keep with h the saved forward activation; TMEM <- tf32(v)
// in place (the next dgrad op's A operand), global copy of delta for the weight-gradient GEMMs through the swizzled
// tile + TMA store, per-32-row column sums (bias gradient) read back column-wise from the same tile.
// The saved activations are read with COALESCED 16-byte global loads (load i of lane l covers row 4i + l/8, columns
// 4 (l%8) .. +3 of the chunk: four full 128-byte row segments per instruction), issued before the accumulator is awaited so
// that their latency hides behind the MMAs of this op, and reach the row-per-thread layout through the warp's swizzled
// staging tile.  (First version: one row per thread, 8 x 16 B -- every instruction touched 32 lines; the op-0 epilogue then
// took 15-17 K clocks per tile, see profiles/r02_chain_trace_small_train.txt.)
struct DgradAux { float4 v[8]; };
__device__ __forceinline__ void chain_dgrad_load_aux(const ChainOp& o, int col0, int64_t row0, int64_t M, int lane, DgradAux& ax) {
  const int q = lane & 7;
  const bool col_ok = o.ep.saved != nullptr && (col0 + q * 4 + 3 < o.N);
  const float* sp = o.ep.saved + col0 + q * 4;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t row = row0 + (lane >> 3) + 4 * i;
    if (col_ok && row < M) asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                                        : "=f"(ax.v[i].x), "=f"(ax.v[i].y), "=f"(ax.v[i].z), "=f"(ax.v[i].w) : "l"(sp + row * o.ep.lds));
    else ax.v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
// column sums over the 32 rows (lanes) of a chunk held one row per lane: transposing butterfly, 31 shuffles, fixed tree
__device__ __forceinline__ float chain_colsum32(float (&v)[32], int lane) {
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
template <int ACT, bool DROP>
__device__ __forceinline__ void chain_dgrad_chunk(const ChainOp& o, const CUtensorMap* tmO, uint32_t taddr, int col0, int64_t tile_row0,
                                                  int64_t M, int quad, int lane, DgradAux& ax, EpiStage& es) {
  uint32_t r[32];
  tc_ld32(taddr, r);
  const uint32_t tile = smem_u32(es.buf[0]);
  if (o.has_out) {
    if (lane == 0) bulk_wait_read<0>();
    __syncwarp();
  }
  const bool staged = o.ep.saved && o.has_out;
  if (staged) {        // coalesced-load layout -> row layout through the staging tile; each thread then reads its own row chunk by chunk
    const int q = lane & 7;
#pragma unroll
    for (int i = 0; i < 8; ++i) { const int rr = (lane >> 3) + 4 * i; sts128(tile + rr * 128 + ((q ^ (rr & 7)) << 4), ax.v[i]); }
    __syncwarp();
  }
  const int64_t lrow = tile_row0 + quad * 32 + lane;
  const int64_t grow = lrow + o.ep.row0;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 h4 = staged ? lds128(row_chunk(tile, lane, q)) : make_float4(0.f, 0.f, 0.f, 0.f);      // (read before this chunk's result overwrites it)
    const float hs[4] = {h4.x, h4.y, h4.z, h4.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float g = __uint_as_float(r[q * 4 + e]);
      float h = hs[e];
      if (DROP) {
        uint32_t w = philox_word((uint64_t)grow * (uint64_t)o.ep.drop_width + (uint64_t)(col0 + q * 4 + e), o.ep.drop_stream, __ldg(o.ep.step), o.ep.seed);
        if ((w >> 8) < o.ep.keep_thr) { g = g / o.ep.keep; h = h * o.ep.keep; } else { g = 0.f; }
      }
      r[q * 4 + e] = __float_as_uint(g * dact_t<ACT>(h));
    }
    if (o.has_out) sts128(row_chunk(tile, lane, q), make_float4(__uint_as_float(r[q * 4]), __uint_as_float(r[q * 4 + 1]),
                                                                __uint_as_float(r[q * 4 + 2]), __uint_as_float(r[q * 4 + 3])));
  }
  float outv[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) { outv[j] = __uint_as_float(r[j]); r[j] = to_tf32(outv[j]); }
  tc_st32(taddr, r);
  if (o.has_out) {
    if (o.ep.colsum_partials) {          // rows past M hold exact zeros (zero-filled operands, zero aux)
      const int64_t row0 = tile_row0 + quad * 32;
      const float cs = chain_colsum32(outv, lane);
      if (row0 < M && col0 + lane < o.N) o.ep.colsum_partials[(row0 >> 5) * o.N + col0 + lane] = cs;
    }
    fence_async_smem();
    __syncwarp();
    if (lane == 0) { tma_store_2d(tmO, es.buf[0], col0, (int)(tile_row0 + quad * 32)); bulk_commit(); }
    es.uses++;
  }
}
template <bool DROP>
__device__ __forceinline__ void chain_dgrad_dispatch(const ChainOp& o, const CUtensorMap* tmO, uint32_t taddr, int col0, int64_t tile_row0,
                                                     int64_t M, int quad, int lane, DgradAux& ax, EpiStage& es) {
  switch (o.ep.act) {
    case MMAE_ACT_RELU: chain_dgrad_chunk<MMAE_ACT_RELU, DROP>(o, tmO, taddr, col0, tile_row0, M, quad, lane, ax, es); break;
    case MMAE_ACT_TANH: chain_dgrad_chunk<MMAE_ACT_TANH, DROP>(o, tmO, taddr, col0, tile_row0, M, quad, lane, ax, es); break;
    case MMAE_ACT_SOFTSIGN: chain_dgrad_chunk<MMAE_ACT_SOFTSIGN, DROP>(o, tmO, taddr, col0, tile_row0, M, quad, lane, ax, es); break;
    case MMAE_ACT_SOFTPLUS: chain_dgrad_chunk<MMAE_ACT_SOFTPLUS, DROP>(o, tmO, taddr, col0, tile_row0, M, quad, lane, ax, es); break;
    default: chain_dgrad_chunk<MMAE_ACT_LINEAR, DROP>(o, tmO, taddr, col0, tile_row0, M, quad, lane, ax, es); break;
  }
}

// final dgrad op: the saved activation tile arrives by TMA (like the loss target), the result overwrites it in place
template <int ACT, bool DROP>
__device__ __forceinline__ void chain_final_dgrad_chunk(const ChainOp& o, uint32_t taddr, int col0, int64_t grow, int lane, uint32_t tile) {
  uint32_t r[32];
  tc_ld32(taddr, r);
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 h4 = lds128(row_chunk(tile, lane, q));
    const float hs[4] = {h4.x, h4.y, h4.z, h4.w};
    float v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float g = __uint_as_float(r[q * 4 + e]);
      float h = hs[e];
      if (DROP) {
        uint32_t w = philox_word((uint64_t)grow * (uint64_t)o.ep.drop_width + (uint64_t)(col0 + q * 4 + e), o.ep.drop_stream, __ldg(o.ep.step), o.ep.seed);
        if ((w >> 8) < o.ep.keep_thr) { g = g / o.ep.keep; h = h * o.ep.keep; } else { g = 0.f; }
      }
      v[e] = g * dact_t<ACT>(h);
    }
    sts128(row_chunk(tile, lane, q), make_float4(v[0], v[1], v[2], v[3]));
  }
}
template <bool DROP>
__device__ __forceinline__ void chain_final_dgrad_dispatch(const ChainOp& o, uint32_t taddr, int col0, int64_t grow, int lane, uint32_t tile) {
  switch (o.ep.act) {
    case MMAE_ACT_RELU: chain_final_dgrad_chunk<MMAE_ACT_RELU, DROP>(o, taddr, col0, grow, lane, tile); break;
    case MMAE_ACT_TANH: chain_final_dgrad_chunk<MMAE_ACT_TANH, DROP>(o, taddr, col0, grow, lane, tile); break;
    case MMAE_ACT_SOFTSIGN: chain_final_dgrad_chunk<MMAE_ACT_SOFTSIGN, DROP>(o, taddr, col0, grow, lane, tile); break;
    case MMAE_ACT_SOFTPLUS: chain_final_dgrad_chunk<MMAE_ACT_SOFTPLUS, DROP>(o, taddr, col0, grow, lane, tile); break;
    default: chain_final_dgrad_chunk<MMAE_ACT_LINEAR, DROP>(o, taddr, col0, grow, lane, tile); break;
  }
}

// ------------------------------------------------------------------ epilogue of the final op, one 32-column chunk
// l = acc + bias; loss += f(l, target); out = dLoss/dl (TRAIN) or decoded_X (PRED).  The target tile was loaded by
// TMA into `tile`; the result overwrites it in place and leaves with a TMA store.  AUX = a target is present.
template <int MODE, int LOSS, bool AUX, bool FULL, bool FILL, bool WL>
__device__ __forceinline__ void chain_final_chunk(const ChainOp& o, uint32_t taddr, int col0, bool row_valid, int lane, uint32_t bias_s,
                                                  uint32_t tile, float& loss_acc, uint32_t miss, uint32_t lut_s) {
  uint32_t r[32];
  tc_ld32(taddr, r);
  float csum = 0.f;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    float4 xq[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) xq[u] = AUX ? lds128(row_chunk(tile, lane, h * 4 + u)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int q = h * 4 + u;
      const float4 b = lds128_ro(bias_s + (col0 + q * 4) * 4);
      const float4 x4 = xq[u];
      const float xs[4] = {x4.x, x4.y, x4.z, x4.w};
      const float bs[4] = {b.x, b.y, b.z, b.w};
      float outv[4];
      uint32_t cm4 = 0u;                                   // modality of the 4 columns, one byte each (warp-uniform)
      if (FILL) asm("ld.shared.u32 %0, [%1];" : "=r"(cm4) : "r"(lut_s + col0 + q * 4));
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float l = __uint_as_float(r[q * 4 + e]) + bs[e], x = xs[e];
        float lossv = 0.f;
        if (LOSS == MMAE_LOSS_SIGMOID_CE) {
          // t = exp(-|l|); sigmoid = 1/(1+t) mirrored for l < 0 with one LOP3 (0.5 + copysign(inv - 0.5, l));
          // softplus(l) = max(l, 0) + log1p(t), log1p on the FMA pipe: two MUFU ops per element, no predicates
          const float t = ex2_ftz(fabsf(l) * -1.4426950408889634f);
          const float inv = rcp_ftz(1.f + t);
          const float hm = inv - 0.5f;
          const float sg = __int_as_float((__float_as_int(hm) & 0x7fffffff) | (__float_as_int(l) & 0x80000000));
          if (WL) lossv = fmaf(-l, x, fmaxf(l, 0.f)) + log1p_unit(t);
          outv[e] = (MODE == EPI_LOSS_TRAIN) ? (sg + (0.5f - x)) : (sg + 0.5f);
        } else if (LOSS == MMAE_LOSS_RMSE) {
          const float d = l - x;
          if (WL) lossv = d * d;
          outv[e] = (MODE == EPI_LOSS_TRAIN) ? d : l;
        } else {
          if (WL) lossv = -x * __logf(l);
          outv[e] = (MODE == EPI_LOSS_TRAIN) ? __fdividef(-x, l) : l;
        }
        if (WL) csum += (FULL || col0 + q * 4 + e < o.N) ? lossv : 0.f;
        if (FILL) outv[e] = ((miss >> ((cm4 >> (8 * e)) & 31u)) & 1u) ? outv[e] : x;
      }
      sts128(row_chunk(tile, lane, q), make_float4(outv[0], outv[1], outv[2], outv[3]));
    }
  }
  if (WL && row_valid) loss_acc += csum;
}

template <int MODE, bool AUX, bool FILL = false>
__device__ __forceinline__ void chain_final_dispatch(const ChainOp& o, uint32_t taddr, int col0, bool row_valid, int lane, uint32_t bias_s,
                                                     uint32_t tile, float& loss_acc, uint32_t miss = 0u, uint32_t lut_s = 0u) {
  const bool full = col0 + 32 <= o.N;          // warp-uniform: no per-column predicate needed
#define CH_FINAL(L) do { \
    if (FILL && !o.ep.loss_partials) { \
      if (full) chain_final_chunk<MODE, L, AUX, true, FILL, false>(o, taddr, col0, row_valid, lane, bias_s, tile, loss_acc, miss, lut_s); \
      else chain_final_chunk<MODE, L, AUX, false, FILL, false>(o, taddr, col0, row_valid, lane, bias_s, tile, loss_acc, miss, lut_s); \
    } else { \
      if (full) chain_final_chunk<MODE, L, AUX, true, FILL, AUX>(o, taddr, col0, row_valid, lane, bias_s, tile, loss_acc, miss, lut_s); \
      else chain_final_chunk<MODE, L, AUX, false, FILL, AUX>(o, taddr, col0, row_valid, lane, bias_s, tile, loss_acc, miss, lut_s); \
    } } while (0)
  switch (o.ep.loss) {
    case MMAE_LOSS_SIGMOID_CE: CH_FINAL(MMAE_LOSS_SIGMOID_CE); break;
    case MMAE_LOSS_RMSE: CH_FINAL(MMAE_LOSS_RMSE); break;
    default: CH_FINAL(MMAE_LOSS_CE); break;
  }
#undef CH_FINAL
}

// ------------------------------------------------------------------ the kernel
// BWD = false: forward chain (bias + activation epilogues, loss / fill-in in the final op);  BWD = true: backward dgrad
// chain (act' epilogues).  Two instantiations so that neither carries the other's registers.
template <bool BWD>
__global__ void __launch_bounds__(CH_THREADS, 1) chain_tc_kernel(const __grid_constant__ ChainParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* xring = smem;
  uint8_t* wring = xring + p.x_stages * CH_XBYTES;
  uint8_t* hid_tiles = xring + CH_POOL_BYTES - CH_HID_BYTES;       // only used (and reserved) when p.hid_tiles
  uint8_t* out_tiles = xring + CH_POOL_BYTES;
  float* bias_s = reinterpret_cast<float*>(out_tiles + CH_OUT_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(bias_s) + CH_BIAS_FLOATS * 4);
  uint64_t* xfull = bars;                               // [CH_MAX_XSTAGES]
  uint64_t* xempty = xfull + CH_MAX_XSTAGES;            // [CH_MAX_XSTAGES]
  uint64_t* wfull = xempty + CH_MAX_XSTAGES;            // [CH_MAX_WSTAGES]
  uint64_t* wempty = wfull + CH_MAX_WSTAGES;            // [CH_MAX_WSTAGES]
  uint64_t* mma_done = wempty + CH_MAX_WSTAGES;         // [CH_MAX_OPS]      accumulator of op i complete
  uint64_t* chunk_done = mma_done + CH_MAX_OPS;         // [CH_MAX_OPS][CH_MAX_CHUNKS]  32 activated columns back in TMEM
  uint64_t* last_done = chunk_done + (CH_MAX_OPS - 1) * CH_MAX_CHUNKS;   // [1]    final op's accumulator drained (non-final ops <= 7)
  uint64_t* aux_bar = last_done + 1;                    // [CH_EPI_WARPS][2]
  uint64_t* miss_ready = aux_bar + CH_EPI_WARPS * 2;    // [2]  missing-block bits of a tile written (scan warp -> output warps)
  uint64_t* miss_free = miss_ready + 2;                 // [2]  ... consumed (output warps -> scan warp)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(miss_free + 2);
  float* epi_red = reinterpret_cast<float*>(tmem_slot + 2);       // [CH_EPI_WARPS]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int last = p.nops - 1;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tmA) : "memory");
    for (int i = 0; i < p.nops; ++i) asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tmB[i]) : "memory");
  }
  if (warp == 2 && lane == 0) {
    for (int s = 0; s < CH_MAX_XSTAGES; ++s) { mbar_init(&xfull[s], 1); mbar_init(&xempty[s], p.scan_miss ? 1 + CH_EPI_WARPS : 1); }
    for (int s = 0; s < CH_MAX_WSTAGES; ++s) { mbar_init(&wfull[s], 1); mbar_init(&wempty[s], 1); }
    for (int i = 0; i < CH_MAX_OPS; ++i) mbar_init(&mma_done[i], 1);
    for (int i = 0; i < (CH_MAX_OPS - 1) * CH_MAX_CHUNKS; ++i) mbar_init(&chunk_done[i], 4);     // one warp per lane quadrant
    mbar_init(last_done, CH_EPI_WARPS);
    for (int i = 0; i < 2; ++i) { mbar_init(&miss_ready[i], CH_EPI_WARPS); mbar_init(&miss_free[i], CH_EPI_WARPS); }
    for (int i = 0; i < CH_EPI_WARPS * 2; ++i) mbar_init(&aux_bar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 3) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(CH_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // zero-padded bias table: bias_s[op.bias_off + col]
  for (int i = 0; i < p.nops; ++i) {
    const ChainOp& o = p.op[i];
    const int w = o.n_chunk * o.n_chunks;
    for (int c = threadIdx.x; c < w; c += CH_THREADS)
      bias_s[o.bias_off + c] = (o.ep.bias && c < o.N) ? __ldg(o.ep.bias + c) : 0.f;
  }
  uint32_t* miss_s = reinterpret_cast<uint32_t*>(bias_s + CH_BIAS_FLOATS - 128 - 256);   // [2][128] missing-block bits per row
  uint8_t* lut = reinterpret_cast<uint8_t*>(bias_s + CH_BIAS_FLOATS - 128);          // column -> modality of the final op (fill-in)
  if (p.op[last].ep.fill_bits)
    for (int c = threadIdx.x; c < 512; c += CH_THREADS) lut[c] = c < p.op[last].N ? p.op[last].ep.fill_col_mod[c] : 0;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== X producer: rows of this CTA's tiles, k-chunk by k-chunk =====================
    if (lane == 0) {
      if (p.stagger_ns) __nanosleep(((blockIdx.x * 61u) % gridDim.x) * p.stagger_ns);      // spread the CTAs' phases (see chain_tc.cuh)
      int stage = 0; uint32_t phase = 0;
      const int K0 = p.op[0].K;
      const uint64_t pol_keep = l2_policy_evict_last();
      for (int t = blockIdx.x; t < p.m_tiles; t += gridDim.x) {
        for (int k = 0; k < K0; k += TC_BK) {
          mbar_wait_parked(&xempty[stage], phase ^ 1);
          mbar_expect_tx(&xfull[stage], CH_XBYTES);
          tma_load_2d_hint(&p.tmA, &xfull[stage], xring + stage * CH_XBYTES, k, t * TC_BM, pol_keep);
          if (++stage == p.x_stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== W producer: every op's weight k-chunks, once per tile (L2-resident) =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      const uint64_t pol_keep = l2_policy_evict_last();
      const int nops = p.nops, w_stages = p.w_stages;
      const uint32_t w_slot = (uint32_t)p.w_slot_bytes;
      const uint32_t wring_s = smem_u32(wring), wfull_s = smem_u32(wfull), wempty_s = smem_u32(wempty);
      int opK[CH_MAX_OPS], opNC[CH_MAX_OPS], opNCs[CH_MAX_OPS];
#pragma unroll
      for (int i = 0; i < CH_MAX_OPS; ++i) { opK[i] = p.op[i].K; opNC[i] = p.op[i].n_chunk; opNCs[i] = p.op[i].n_chunks; }
      for (int t = blockIdx.x; t < p.m_tiles; t += gridDim.x) {
#pragma unroll
        for (int i = 0; i < CH_MAX_OPS; ++i) {
          if (i < nops) {
            const uint32_t bytes = (uint32_t)opNC[i] * TC_BK * 4;
            for (int nc = 0; nc < opNCs[i]; ++nc) {
              for (int k = 0; k < opK[i]; k += TC_BK) {
                mbar_wait_s(wempty_s + stage * 8u, phase ^ 1);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(wfull_s + stage * 8u), "r"(bytes) : "memory");
                asm volatile(
                    "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
                    ::"r"(wring_s + stage * w_slot), "l"(&p.tmB[i]), "r"(wfull_s + stage * 8u), "r"(k), "r"(nc * opNC[i]), "l"(pol_keep) : "memory");
                if (++stage == w_stages) { stage = 0; phase ^= 1; }
              }
            }
          }
        }
      }
    }
  } else if (warp == 2) {
    // ===================== MMA issuer =====================
    // One lane runs the whole loop (the rest of the warp parks at the final barrier): its instruction stream is
    // serial and shares an issue port with four busy epilogue warps, so every per-chunk instruction counts --
    // per-op fields are hoisted out of the constant bank, descriptors are built by adding to a precomputed base.
    if (lane == 0) {
      const uint32_t w_slot = (uint32_t)p.w_slot_bytes;
      const int w_stages = p.w_stages, x_stages = p.x_stages, nops = p.nops;
      const uint32_t wring_s = smem_u32(wring), xring_s = smem_u32(xring);
      const uint64_t desc_hi = make_smem_desc(0, 16, 1024);            // K-major SW128: LBO 16 B, SBO 1024 B; address field added below
      const uint32_t xfull_s = smem_u32(xfull), xempty_s = smem_u32(xempty), wfull_s = smem_u32(wfull), wempty_s = smem_u32(wempty);
      int xs = 0, ws = 0; uint32_t xph = 0, wph = 0;
      uint32_t it = 0;
      for (int t = blockIdx.x; t < p.m_tiles; t += gridDim.x, ++it) {
        const uint32_t par = it & 1;
        for (int i = 0; i < nops; ++i) {
          const int K = p.op[i].K, a_tmem = p.op[i].a_tmem, n_chunk = p.op[i].n_chunk, n_chunks = p.op[i].n_chunks;
          const uint32_t d_col = (uint32_t)p.op[i].d_col, a_col = (uint32_t)p.op[i].a_col;
          if (i == last && it > 0) { mbar_wait_parked(last_done, par ^ 1); tc_fence_after(); }   // previous tile's result drained
          const uint32_t idesc = idesc_tf32_rt(TC_BM, n_chunk);
          const uint32_t cd_s = smem_u32(&chunk_done[(i > 0 ? i - 1 : 0) * CH_MAX_CHUNKS]);
#ifdef MMAE_CHAIN_TRACE_WAITS
          const bool tracing = p.trace && blockIdx.x == 0 && it < 64;
          long long wait_a = 0, wait_w = 0, wait_x = 0;
          if (tracing) p.trace[(it * CH_MAX_OPS + i) * 4 + 0] = clock64();
#else
          if (p.trace && blockIdx.x == 0 && it < 64) p.trace[(it * CH_MAX_OPS + i) * 4 + 0] = clock64();
#endif
          for (int nc = 0; nc < n_chunks; ++nc) {
            const uint32_t tmem_d = tmem_base + d_col + (uint32_t)(nc * n_chunk);
            uint32_t accumulate = 0;
            for (int k = 0; k < K; k += TC_BK) {
#ifdef MMAE_CHAIN_TRACE_WAITS      // (compile-time: the single-lane issuer must stay lean, see the note above)
              if (tracing) {      // debug timeline: where the issuer waits (activated A chunk / weight chunk / X chunk)
                const long long c0 = clock64();
                if (a_tmem && nc == 0) mbar_wait_s(cd_s + (uint32_t)(k >> 5) * 8u, par);
                const long long c1 = clock64();
                mbar_wait_s(wfull_s + ws * 8u, wph);
                const long long c2 = clock64();
                if (!a_tmem) mbar_wait_s(xfull_s + xs * 8u, xph);
                const long long c3 = clock64();
                wait_a += c1 - c0; wait_w += c2 - c1; wait_x += c3 - c2;
              } else
#endif
              {
              if (a_tmem && nc == 0) mbar_wait_s(cd_s + (uint32_t)(k >> 5) * 8u, par);    // A columns [k, k+32) written
              mbar_wait_s(wfull_s + ws * 8u, wph);
              if (!a_tmem) mbar_wait_s(xfull_s + xs * 8u, xph);
              }
              tc_fence_after();
              const uint64_t bdesc = desc_hi | (uint64_t)(((wring_s + ws * w_slot) >> 4) & 0x3FFF);
              const int rem = K - k;
              if (a_tmem) {
                const uint32_t ta = tmem_base + a_col + (uint32_t)k;
                tc_mma_tf32_ts(tmem_d, ta, bdesc, idesc, accumulate);
                if (rem > 8) tc_mma_tf32_ts(tmem_d, ta + 8, bdesc + 2, idesc, 1);
                if (rem > 16) tc_mma_tf32_ts(tmem_d, ta + 16, bdesc + 4, idesc, 1);
                if (rem > 24) tc_mma_tf32_ts(tmem_d, ta + 24, bdesc + 6, idesc, 1);
              } else {
                const uint64_t adesc = desc_hi | (uint64_t)(((xring_s + xs * CH_XBYTES) >> 4) & 0x3FFF);
                tc_mma_tf32(tmem_d, adesc, bdesc, idesc, accumulate);
                if (rem > 8) tc_mma_tf32(tmem_d, adesc + 2, bdesc + 2, idesc, 1);
                if (rem > 16) tc_mma_tf32(tmem_d, adesc + 4, bdesc + 4, idesc, 1);
                if (rem > 24) tc_mma_tf32(tmem_d, adesc + 6, bdesc + 6, idesc, 1);
                tc_commit_s(xempty_s + xs * 8u);
                if (++xs == x_stages) { xs = 0; xph ^= 1; }
              }
              accumulate = 1;
              tc_commit_s(wempty_s + ws * 8u);
              if (++ws == w_stages) { ws = 0; wph ^= 1; }
            }
          }
          tc_commit(&mma_done[i]);
#ifdef MMAE_CHAIN_TRACE_WAITS
          if (tracing) {
            p.trace[(it * CH_MAX_OPS + i) * 4 + 1] = clock64();
            long long* w = p.trace + 64 * CH_MAX_OPS * 4 + (it * CH_MAX_OPS + i) * 4;
            w[0] = wait_a; w[1] = wait_w; w[2] = wait_x;
          }
#else
          if (p.trace && blockIdx.x == 0 && it < 64) p.trace[(it * CH_MAX_OPS + i) * 4 + 1] = clock64();
#endif
        }
      }
    }
  } else if (warp >= CH_HID_WARP0 && warp < CH_OUT_WARP0) {
    // ===================== hidden epilogue warps: non-final ops, activated values back into TMEM =====================
    const int ew = warp - CH_HID_WARP0;
    const int quad = warp & 3;
    const int half = ew >> 2;
    EpiStage es;
    es.buf[0] = es.buf[1] = hid_tiles + ew * CH_EPI_TILE;
    es.aux_bar[0] = es.aux_bar[1] = nullptr; es.aux_phase[0] = es.aux_phase[1] = 0; es.uses = 0;
    uint32_t it = 0;
    int sxs = 0; uint32_t sxph = 0;
    for (int t = blockIdx.x; t < p.m_tiles; t += gridDim.x, ++it) {
      const uint32_t par = it & 1;
      const int64_t tile_row0 = (int64_t)t * TC_BM;
      if (!BWD && p.scan_miss) {
        // fill-in: missing-block detection (sum == -width, data_funcs.py:366-381) from the X tile while it sits in the
        // ring.  These warps are idle until the first layer's accumulator is complete; warp w owns rows [16w, 16w+16),
        // a lane pair splits the 32 columns of a k-chunk.  Modalities are contiguous, increasing column ranges, so one
        // running sum per row covers a block that spans several k-chunks.
        const int K0 = p.op[0].K, x_stages = p.x_stages, num_mod = p.num_mod;
        const uint32_t xring_s = smem_u32(xring);
        const int r = ew * 16 + (lane & 15), hq = (lane >> 4) * 4;
        float carry = 0.f; uint32_t bits = 0u; int m = 0;
        for (int k0 = 0; k0 < K0; k0 += TC_BK) {
          mbar_wait_parked(&xfull[sxs], sxph);
          const uint32_t base = xring_s + sxs * CH_XBYTES + r * 128;
          const int kend = k0 + TC_BK;
          float4 f[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) f[q] = lds128(base + (((hq + q) ^ (r & 7)) << 4));
          __syncwarp();
          if (lane == 0) mbar_arrive(&xempty[sxs]);          // with the MMA commit: 1 + 8 arrivals free the stage
          if (++sxs == x_stages) { sxs = 0; sxph ^= 1; }
          while (m < num_mod) {
            const int s = __ldg(p.starts + m), e1 = __ldg(p.starts + m + 1);
            if (s >= kend) break;
            float part = 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int c = k0 + (hq + q) * 4;
              part += (c >= s && c < e1) ? f[q].x : 0.f;
              part += (c + 1 >= s && c + 1 < e1) ? f[q].y : 0.f;
              part += (c + 2 >= s && c + 2 < e1) ? f[q].z : 0.f;
              part += (c + 3 >= s && c + 3 < e1) ? f[q].w : 0.f;
            }
            part += __shfl_xor_sync(0xffffffffu, part, 16);
            if (e1 <= kend) { if (carry + part == -(float)(e1 - s)) bits |= 1u << m; carry = 0.f; ++m; }
            else { carry += part; break; }
          }
        }
        if (it >= 2) mbar_wait_parked(&miss_free[it & 1], ((it >> 1) - 1) & 1);      // the tile that last used this slot is drained
        if (lane < 16) miss_s[(it & 1) * 128 + r] = bits;
        __syncwarp();
        if (lane == 0) mbar_arrive(&miss_ready[it & 1]);
      }
      for (int i = 0; i < last; ++i) {
        const ChainOp& o = p.op[i];
        DgradAux dax;
        if (BWD && half < o.n_chunk / 32) chain_dgrad_load_aux(o, half * 32, tile_row0 + quad * 32, p.M, lane, dax);
        mbar_wait_parked(&mma_done[i], par);
        tc_fence_after();
        if (p.trace && blockIdx.x == 0 && ew == 0 && lane == 0 && it < 64) p.trace[(it * CH_MAX_OPS + i) * 4 + 2] = clock64();
        const int chunks = o.n_chunk / 32;
        const uint32_t tbase = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)o.d_col;
        for (int ch = half; ch < chunks; ch += 2) {
          if (BWD) {
            if (ch != half) chain_dgrad_load_aux(o, ch * 32, tile_row0 + quad * 32, p.M, lane, dax);     // (the first chunk's loads were issued before the wait)
            if (o.ep.keep < 1.f) chain_dgrad_dispatch<true>(o, &p.tmO[i], tbase + ch * 32, ch * 32, tile_row0, p.M, quad, lane, dax, es);
            else chain_dgrad_dispatch<false>(o, &p.tmO[i], tbase + ch * 32, ch * 32, tile_row0, p.M, quad, lane, dax, es);
          } else if (o.ep.keep < 1.f) chain_act_dispatch<true>(o, &p.tmO[i], tbase + ch * 32, ch * 32, tile_row0, quad, lane, smem_u32(bias_s + o.bias_off), es);
          else chain_act_dispatch<false>(o, &p.tmO[i], tbase + ch * 32, ch * 32, tile_row0, quad, lane, smem_u32(bias_s + o.bias_off), es);
          tc_wait_st();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&chunk_done[i * CH_MAX_CHUNKS + ch]);
        }
        if (p.trace && blockIdx.x == 0 && ew == 0 && lane == 0 && it < 64) p.trace[(it * CH_MAX_OPS + i) * 4 + 3] = clock64();
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");     // every saved activation has landed
  } else if (warp >= CH_OUT_WARP0) {
    // ===================== output epilogue warps: the final op (loss, delta_L / decoded_X) =====================
    const int ew = warp - CH_OUT_WARP0;
    const int quad = warp & 3;
    const int half = ew >> 2;
    EpiStage es;
    es.buf[0] = out_tiles + ew * 2 * CH_EPI_TILE; es.buf[1] = es.buf[0] + CH_EPI_TILE;
    es.aux_bar[0] = &aux_bar[ew * 2]; es.aux_bar[1] = &aux_bar[ew * 2 + 1];
    es.aux_phase[0] = es.aux_phase[1] = 0; es.uses = 0;
    float loss_acc = 0.f;
    const ChainOp& lo = p.op[last];
    const bool has_aux = lo.ep.target != nullptr;
    const int lchunks = lo.n_chunks * lo.n_chunk / 32;
    const uint32_t tbase = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)lo.d_col;
    const uint32_t lbias = smem_u32(bias_s + lo.bias_off);
    const uint64_t pol_stream = l2_policy_evict_first();
    uint32_t it = 0;
    for (int t = blockIdx.x; t < p.m_tiles; t += gridDim.x, ++it) {
      const uint32_t par = it & 1;
      const int row0 = (int)((int64_t)t * TC_BM + quad * 32);
      // the target tile of my first chunk travels while the chain of this tile is still running
      if (has_aux && half < lchunks && half * 32 < lo.N) {
        const int b = es.uses & 1;
        if (lane == 0) {
          bulk_wait_read<1>();          // the store that last used this buffer is done reading it
          mbar_expect_tx(es.aux_bar[b], CH_EPI_TILE);
          tma_load_2d_hint(&p.tmT, es.aux_bar[b], es.buf[b], half * 32, row0, pol_stream);
        }
        __syncwarp();
      }
      mbar_wait_parked(&mma_done[last], par);
      tc_fence_after();
      if (p.trace && blockIdx.x == 0 && ew == 0 && lane == 0 && it < 64) p.trace[(it * CH_MAX_OPS + last) * 4 + 2] = clock64();
      const bool row_valid = (int64_t)row0 + lane < p.M;
      uint32_t miss = 0u;
      if (!BWD && p.scan_miss) { mbar_wait_parked(&miss_ready[it & 1], (it >> 1) & 1); miss = miss_s[(it & 1) * 128 + quad * 32 + lane]; }
      else if (!BWD && lo.ep.fill_bits && row_valid) miss = __ldg(lo.ep.fill_bits + row0 + lane);
      for (int ch = half; ch < lchunks; ch += 2) {
        if (ch * 32 >= lo.N) break;                                   // padding columns only
        const int b = es.uses & 1;
        const int nxt = ch + 2;
        if (has_aux && nxt < lchunks && nxt * 32 < lo.N) {            // prefetch the next chunk's target tile
          if (lane == 0) {
            bulk_wait_read<0>();
            mbar_expect_tx(es.aux_bar[b ^ 1], CH_EPI_TILE);
            tma_load_2d_hint(&p.tmT, es.aux_bar[b ^ 1], es.buf[b ^ 1], nxt * 32, row0, pol_stream);
          }
          __syncwarp();
        } else if (!has_aux) {
          if (lane == 0) bulk_wait_read<1>();
          __syncwarp();
        }
        if (has_aux) { mbar_wait_parked(es.aux_bar[b], es.aux_phase[b]); es.aux_phase[b] ^= 1; }
        uint8_t* tile_p = es.buf[b];
        const uint32_t tile = smem_u32(tile_p);
        if (BWD) {
          const int64_t grow = (int64_t)row0 + lane + lo.ep.row0;
          if (lo.ep.keep < 1.f) chain_final_dgrad_dispatch<true>(lo, tbase + ch * 32, ch * 32, grow, lane, tile);
          else chain_final_dgrad_dispatch<false>(lo, tbase + ch * 32, ch * 32, grow, lane, tile);
        } else if (lo.ep.mode == EPI_LOSS_TRAIN) {
          if (has_aux) chain_final_dispatch<EPI_LOSS_TRAIN, true>(lo, tbase + ch * 32, ch * 32, row_valid, lane, lbias, tile, loss_acc);
          else chain_final_dispatch<EPI_LOSS_TRAIN, false>(lo, tbase + ch * 32, ch * 32, row_valid, lane, lbias, tile, loss_acc);
        } else if (lo.ep.fill_bits && has_aux) {
          chain_final_dispatch<EPI_LOSS_PRED, true, true>(lo, tbase + ch * 32, ch * 32, row_valid, lane, lbias, tile, loss_acc, miss, smem_u32(lut));
        } else {
          if (has_aux) chain_final_dispatch<EPI_LOSS_PRED, true>(lo, tbase + ch * 32, ch * 32, row_valid, lane, lbias, tile, loss_acc);
          else chain_final_dispatch<EPI_LOSS_PRED, false>(lo, tbase + ch * 32, ch * 32, row_valid, lane, lbias, tile, loss_acc);
        }
        __syncwarp();
        if (lo.ep.colsum_partials && row0 < p.M) {
          // bias gradient: column sums of this warp's 32 rows, read back column-wise from the swizzled tile
          const int col = ch * 32 + lane;
          const int nrows = (int)min((int64_t)32, p.M - row0);
          float cs = 0.f;
          for (int r = 0; r < nrows; ++r)
            cs += lds32(tile + r * 128 + (((lane >> 2) ^ (r & 7)) << 4) + (lane & 3) * 4);
          if (col < lo.N) lo.ep.colsum_partials[(int64_t)(row0 >> 5) * lo.N + col] = cs;
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) { tma_store_2d_hint(&p.tmO[last], tile_p, ch * 32, row0, pol_stream); bulk_commit(); }
        es.uses++;
      }
      tc_fence_before();
      __syncwarp();
      if (p.trace && blockIdx.x == 0 && ew == 0 && lane == 0 && it < 64) p.trace[(it * CH_MAX_OPS + last) * 4 + 3] = clock64();
      if (lane == 0) { mbar_arrive(last_done); if (p.scan_miss) mbar_arrive(&miss_free[it & 1]); }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");     // every output tile has landed
    if (lo.ep.loss_partials) {
      float w = warp_sum(loss_acc);
      if (lane == 0) epi_red[ew] = w;
      asm volatile("bar.sync 1, 256;" ::: "memory");     // the 8 output warps only
      if (ew == 0 && lane == 0) {
        float sacc = 0.f;
#pragma unroll
        for (int i = 0; i < CH_EPI_WARPS; ++i) sacc += epi_red[i];
        lo.ep.loss_partials[blockIdx.x] = sacc;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 3) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(CH_TMEM_COLS) : "memory");
  }
}

cudaError_t chain_launch(const ChainParams& p, int grid, cudaStream_t st) {
  static bool configured_dev[64] = {};        // the attribute is per device: one flag per device ordinal
  int dev_ = 0; cudaGetDevice(&dev_);
  bool& configured = configured_dev[dev_ & 63];
  if (!configured || dev_ >= 64) {      // (ordinals past the table are configured on every launch instead of aliasing)
    cudaError_t e = cudaFuncSetAttribute(chain_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, CH_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(chain_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, CH_SMEM);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  if (p.op[0].ep.mode == EPI_DGRAD) chain_tc_kernel<true><<<grid, CH_THREADS, CH_SMEM, st>>>(p);
  else chain_tc_kernel<false><<<grid, CH_THREADS, CH_SMEM, st>>>(p);
  return cudaGetLastError();
}

}  // namespace mmae
