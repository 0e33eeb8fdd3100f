// One-CTA tcgen05 GEMM used by the tests to pin the shared-memory descriptor conventions (K-major and
// MN-major no-swizzle operands, instruction descriptor, TMEM accumulator layout) that the fused kernels
// rely on:  D[128, N] = A[128, K] . B[N, K]^T with fp16 operands and fp32 accumulation.
#pragma once
#include "tc_ptx.cuh"

namespace sisua {
namespace tc {

// A: logical [128][K] row-major fp32 in global, B: logical [N][K] row-major fp32, D: [128][N] fp32.
// a_mn / b_mn select how the operand tile is laid out in shared memory and described to the MMA.
__global__ void __launch_bounds__(128) tc_selftest_kernel(const float* __restrict__ A, const float* __restrict__ Bm,
                                                          float* __restrict__ D, int N, int K, int a_mn, int b_mn) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int M = 128;
  uint8_t* sA = smem;
  uint8_t* sB = smem + (size_t)M * K * 2;
  const int t = threadIdx.x;
  // physical tiles: K-major -> T[rows = m][cols = k]; MN-major -> T[rows = k][cols = m]
  const int A_RS = 128, A_CS = a_mn ? (K / 8) * 128 : (M / 8) * 128;
  const int B_RS = 128, B_CS = b_mn ? (K / 8) * 128 : (N / 8) * 128;
  for (int i = t; i < M * K; i += blockDim.x) {
    int m = i / K, k = i % K;
    uint32_t off = a_mn ? tile_off(k, m, A_RS, A_CS) : tile_off(m, k, A_RS, A_CS);
    *reinterpret_cast<__half*>(sA + off) = __float2half_rn(A[i]);
  }
  for (int i = t; i < N * K; i += blockDim.x) {
    int n = i / K, k = i % K;
    uint32_t off = b_mn ? tile_off(k, n, B_RS, B_CS) : tile_off(n, k, B_RS, B_CS);
    *reinterpret_cast<__half*>(sB + off) = __float2half_rn(Bm[i]);
  }
  if (t == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (t < 32) tmem_alloc(&tmem_base_s, 256);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (t == 0) {
    const uint32_t idesc = make_idesc_f16(M, N, a_mn, b_mn);
    // K-major: SBO = RS (between 8-row groups of the M/N index), LBO = CS (between 16-byte K chunks);
    // MN-major: SBO = CS (between 8-element groups of the M/N index), LBO = RS (between 8-row K groups).
    const uint32_t a_lbo = a_mn ? A_RS : A_CS, a_sbo = a_mn ? A_CS : A_RS;
    const uint32_t b_lbo = b_mn ? B_RS : B_CS, b_sbo = b_mn ? B_CS : B_RS;
    for (int ks = 0; ks < K / 16; ++ks) {
      uint64_t ad = make_smem_desc(smem_u32(sA) + ks * 2 * a_lbo, a_lbo, a_sbo);
      uint64_t bd = make_smem_desc(smem_u32(sB) + ks * 2 * b_lbo, b_lbo, b_sbo);
      umma_f16(tmem, ad, bd, idesc, ks > 0 ? 1u : 0u);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  const int warp = t >> 5;
  for (int c0 = 0; c0 < N; c0 += 16) {
    float v[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) D[(size_t)t * N + c0 + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (t < 32) tmem_dealloc(tmem, 256);
}

}  // namespace tc
}  // namespace sisua
