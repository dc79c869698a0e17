// Single-tile tcgen05 tf32 probe: C[128,64] = A[128,32] * B[64,32]^T, both K-major, SW128.
// Sweeps descriptor fields given on the command line; dumps smem after TMA to verify the load.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include "../../multimodalautoencoder_b200/csrc/gemm_tc_kernel.cuh"
using namespace mmae;

struct PP { CUtensorMap ta, tb; float* C; float* dumpA; uint32_t lbo, sbo, idesc; int kstep; int amn; uint32_t alt; };

__global__ void __launch_bounds__(128, 1) probe(const __grid_constant__ PP p) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bar = (uint64_t*)(smem + 16384 + 8192);
  uint64_t* mbar2 = bar + 1;
  uint32_t* slot = (uint32_t*)(bar + 2);
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(mbar2, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "n"(64) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  uint32_t tm = *slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, 16384 + 8192);
    if (!p.amn) tma_load_2d(&p.ta, bar, smem, 0, 0);
    else for (int c = 0; c < 4; ++c) tma_load_2d(&p.ta, bar, smem + c * 4096, c * 32, 0);
    tma_load_2d(&p.tb, bar, smem + 16384, 0, 0);
  }
  mbar_wait(bar, 0);
  // dump A tile as it sits in smem
  for (int i = threadIdx.x; i < 4096; i += 128) p.dumpA[i] = ((float*)smem)[i];
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    uint32_t sa = smem_u32(smem), sb = sa + 16384;
    for (int kk = 0; kk < 4; ++kk) {
      uint64_t ad = !p.amn ? make_smem_desc(sa + kk * 32, 16, 1024) : make_smem_desc(sa + kk * p.kstep, p.lbo, p.sbo, p.alt);
      uint64_t bd = make_smem_desc(sb + kk * 32, 16, 1024);
      tc_mma_tf32(tm, ad, bd, p.idesc, kk > 0);
    }
    tc_commit(mbar2);
  }
  mbar_wait(mbar2, 0);
  tc_fence_after();
  uint32_t r[32];
  for (int ch = 0; ch < 2; ++ch) {
    tc_ld32(tm + ((uint32_t)(warp * 32) << 16) + ch * 32, r);
    for (int j = 0; j < 32; ++j) p.C[(warp * 32 + lane) * 64 + ch * 32 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "n"(64) : "memory");
}

int main(int argc, char** argv) {
  const int M = 128, N = 64, K = 32;
  float *hA = (float*)malloc(M * K * 4), *hB = (float*)malloc(N * K * 4), *hC = (float*)malloc(M * N * 4), *hD = (float*)malloc(16384);
  for (int i = 0; i < M * K; ++i) hA[i] = (float)((i * 7 + 3) % 13) - 6.f;
  for (int i = 0; i < N * K; ++i) hB[i] = (float)((i * 5 + 1) % 11) - 5.f;
  float *dA, *dB, *dC, *dD;
  cudaMalloc(&dA, M * K * 4); cudaMalloc(&dB, N * K * 4); cudaMalloc(&dC, M * N * 4); cudaMalloc(&dD, 16384);
  cudaMemcpy(dA, hA, M * K * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, N * K * 4, cudaMemcpyHostToDevice);
  PP p;
  if (!make_tmap(&p.ta, dA, M, K, K, 32, 128) || !make_tmap(&p.tb, dB, N, K, K, 32, 64)) { printf("tmap encode failed\n"); return 1; }
  p.C = dC; p.dumpA = dD;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
  // A stored MN-major [K, M]
  float* hAt = (float*)malloc(M * K * 4);
  for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) hAt[k * M + m] = hA[m * K + k];
  float* dAt; cudaMalloc(&dAt, M * K * 4); cudaMemcpy(dAt, hAt, M * K * 4, cudaMemcpyHostToDevice);
  struct V { int amn; uint32_t alt, lbo, sbo; int kstep; bool sw32; const char* name; } vs[] = {
    {0, 2, 16, 1024, 32, false, "K-major reference"},
    {1, 1, 4096, 512, 1024, true, "MN sw128_32B lbo4096 sbo512 k1024"},
    {1, 1, 512, 4096, 1024, true, "MN sw128_32B lbo512 sbo4096 (swapped)"},
    {1, 1, 4096, 1024, 1024, true, "MN sw128_32B lbo4096 sbo1024"},
    {1, 2, 4096, 1024, 1024, false, "MN sw128 (16B atom) lbo4096 sbo1024"},
    {1, 2, 4096, 1024, 1024, true, "MN desc sw128, tma 32B"},
    {1, 1, 4096, 512, 1024, false, "MN desc 32B, tma sw128"}};
  for (auto& v : vs) {
    if (v.amn) { if (!make_tmap(&p.ta, dAt, K, M, M, 32, 32, v.sw32)) { printf("tmap fail\n"); return 1; } }
    else if (!make_tmap(&p.ta, dA, M, K, K, 32, 128)) return 1;
    p.lbo = v.lbo; p.sbo = v.sbo; p.kstep = v.kstep; p.amn = v.amn; p.alt = v.alt;
    p.idesc = make_idesc_tf32(128, 64, v.amn != 0, false);
    cudaMemset(dC, 0xff, M * N * 4);
    probe<<<1, 128, 40000>>>(p);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: CUDA error %s\n", v.name, cudaGetErrorString(e)); return 2; }
    cudaMemcpy(hC, dC, M * N * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0; int nz = 0;
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
      double sacc = 0; for (int k = 0; k < K; ++k) sacc += (double)hA[m * K + k] * hB[n * K + k];
      maxerr = fmax(maxerr, fabs(sacc - hC[m * N + n])); maxref = fmax(maxref, fabs(sacc)); nz += hC[m * N + n] != 0.f;
    }
    printf("%-44s idesc=%08x maxerr=%g maxref=%g nonzero=%d\n", v.name, p.idesc, maxerr, maxref, nz);
  }
  return 0;
}
