// gemm_mode 0: the genes x hidden contractions as plain fp32 CUDA-core GEMMs with the decoder
// output materialised in HBM, followed by a row kernel for the count likelihood.  This is the
// straightforward op-by-op arrangement (what the reference's TF graph does: SURVEY.md section 3.1)
// kept as the on-GPU cross-check for the fused tcgen05 kernels; it is exact fp32.
#pragma once
#include "device_math.cuh"

namespace sisua {

enum LoadOp : int { LOAD_NONE = 0, LOAD_LOG1P = 1, LOAD_RAW_DROP = 2 };   // 1, 2: the operand is the count matrix (input dropout applies)

// C[M,N] (+)= op_a(A)(m,k) * op_b(B)(k,n);  A(m,k) = A[m*a_rs + k*a_cs], B(k,n) = B[k*b_rs + n*b_cs].
// gridDim.z splits K; with splits > 1 or accumulate the result is added atomically into C.
constexpr int kGemmBM = 64, kGemmBN = 64, kGemmBK = 16;

template <int AOP, int BOP>
__global__ void __launch_bounds__(256) sgemm_kernel(
    const float* __restrict__ A, long long a_rs, long long a_cs, const float* __restrict__ Bm, long long b_rs,
    long long b_cs, float* __restrict__ Cm, long long ldc, const float* __restrict__ bias, int M, int N, int K,
    int k_chunk, int atomic_out, DropSpec drop) {
  __shared__ float As[kGemmBK][kGemmBM + 4];
  __shared__ float Bs[kGemmBK][kGemmBN + 4];
  const int t = threadIdx.x;
  const int m0 = blockIdx.y * kGemmBM, n0 = blockIdx.x * kGemmBN;
  const int kbeg = blockIdx.z * k_chunk, kend = min(K, kbeg + k_chunk);
  const int tm = (t / 16) * 4, tn = (t % 16) * 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const bool a_kfast = (a_cs == 1), b_kfast = (b_rs == 1);
  for (int k0 = kbeg; k0 < kend; k0 += kGemmBK) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int m, k;
      if (a_kfast) { k = t % 16; m = t / 16 + 16 * j; } else { m = t % 64; k = t / 64 + 4 * j; }
      float v = 0.f;
      if (m0 + m < M && k0 + k < kend) {
        v = A[(long long)(m0 + m) * a_rs + (long long)(k0 + k) * a_cs];
        if (AOP == LOAD_LOG1P) v = log1pf(v);
        if (AOP != LOAD_NONE) v *= dropout_mult(drop, (uint32_t)(m0 + m), (uint32_t)(k0 + k));
      }
      As[k][m] = v;
      int n, kb;
      if (b_kfast) { kb = t % 16; n = t / 16 + 16 * j; } else { n = t % 64; kb = t / 64 + 4 * j; }
      float w = 0.f;
      if (n0 + n < N && k0 + kb < kend) {
        w = Bm[(long long)(k0 + kb) * b_rs + (long long)(n0 + n) * b_cs];
        if (BOP == LOAD_LOG1P) w = log1pf(w);
        if (BOP != LOAD_NONE) w *= dropout_mult(drop, (uint32_t)(k0 + kb), (uint32_t)(n0 + n));
      }
      Bs[kb][n] = w;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kGemmBK; ++k) {
      float4 av = *reinterpret_cast<const float4*>(&As[k][tm]);
      float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tn]);
      float a4[4] = {av.x, av.y, av.z, av.w}, b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a4[i], b4[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + tm + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tn + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (bias && blockIdx.z == 0) v += bias[n];
      float* p = Cm + (long long)m * ldc + n;
      if (atomic_out) atomicAdd(p, v); else *p = v;
    }
  }
}

// column sums of a [R, N] matrix, atomically added into out[N] (output-bias gradient)
__global__ void __launch_bounds__(256) col_sum_kernel(const float* __restrict__ A, long long lda, int R, int N,
                                                      int rows_per_block, float* __restrict__ out) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  int r0 = blockIdx.y * rows_per_block, r1 = min(R, r0 + rows_per_block);
  float s = 0.f;
  for (int r = r0; r < r1; ++r) s += A[(long long)r * lda + n];
  atomicAdd(&out[n], s);
}

// ------------------------------------------------------------------------------------------
// count likelihood over one materialised decoder-output row: OUT[r] = (a[G] | b[G] | l[G]).
// One CTA per row.  Train: OUT row is overwritten by d loss / d(a, b, l).
// Rows a7-a9 of SURVEY.md section 8a; scVI parameterisation sisua/models/scvi.py:117-138.
// ------------------------------------------------------------------------------------------
struct CountRowArgs {
  float* OUT; long long ldo;      // [R, nheads*G]
  const float* x;                 // [B, G]
  const float* lib;               // [R] sampled log-library (scVI) or null
  float* llk_x;                   // [R]
  float* dlib;                    // [R] d loss / d lib (scVI train) or null
  float* out_mean; float* out_disp; float* out_pi;   // [R, G] each, nullable (inference)
  int R, B, G;
  int scvi, zero_inflated, train, mean_act, disp_act, reapply;
  int tfp;                        // 'zinb' / 'nb': head 0 = log total_count, head 1 = logits (TFP NegativeBinomial)
  float upstream;                 // d loss / d llk_x = -1 / R
  float clip_library;
};

template <bool ZI>
__global__ void __launch_bounds__(256) count_row_kernel(CountRowArgs a) {
  extern __shared__ float s_cache[];   // scVI: clipped softmax scale of the row, [G]
  __shared__ float scratch[33];
  const int r = blockIdx.x, b = r % a.B, G = a.G;
  float* row = a.OUT + (long long)r * a.ldo;
  const float* xr = a.x + (size_t)b * G;
  float eL = 1.f, row_max = 0.f, inv_sum = 1.f, lib_raw = 0.f;
  bool lib_open = false;
  if (a.scvi) {
    float mx = -INFINITY;
    for (int g = threadIdx.x; g < G; g += blockDim.x) mx = fmaxf(mx, row[g]);
    row_max = block_max(mx, scratch);
    float se = 0.f;
    for (int g = threadIdx.x; g < G; g += blockDim.x) se += expf(row[g] - row_max);
    inv_sum = 1.f / block_sum(se, scratch);
    lib_raw = a.lib[r];
    float lc = fminf(fmaxf(lib_raw, 0.f), a.clip_library);
    lib_open = (lib_raw >= 0.f && lib_raw <= a.clip_library);
    eL = expf(lc);
  }
  float llk = 0.f, dot = 0.f, dl = 0.f;
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    float ra = row[g], rb = row[G + g];
    float pi = ZI ? row[2 * G + g] : 0.f;
    float mu, dmu_da, th, dth_db, s = 0.f, s_raw = 0.f, gate = 1.f;
    if (a.scvi) {
      s_raw = expf(ra - row_max) * inv_sum;
      s = fminf(fmaxf(s_raw, 1e-7f), 1.f - 1e-7f);
      gate = (s_raw >= 1e-7f && s_raw <= 1.f - 1e-7f) ? 1.f : 0.f;
      mu = eL * s; dmu_da = 1.f;
      th = expf(rb); dth_db = th;
      if (a.reapply) {   // Q2 literal reading: activations applied again on the positive params
        float v, dv;
        activation(a.mean_act, mu, v, dv); mu = v; dmu_da = dv;
        activation(a.disp_act, th, v, dv); th = v; dth_db *= dv;
      }
      s_cache[g] = s_raw;
    } else if (a.tfp) {
      th = expf(ra); mu = expf(ra + rb); dmu_da = 1.f; dth_db = 1.f;     // chain rule below
    } else {
      activation(a.mean_act, ra, mu, dmu_da);
      activation(a.disp_act, rb, th, dth_db);
    }
    CountGrad cg;
    float l = a.tfp ? (a.train ? count_llk<ZI, true, true>(xr[g], mu, th, pi, cg) : count_llk<ZI, false, true>(xr[g], mu, th, pi, cg))
                    : (a.train ? count_llk<ZI, true>(xr[g], mu, th, pi, cg) : count_llk<ZI, false>(xr[g], mu, th, pi, cg));
    llk += l;
    if (a.train && a.tfp) {      // d / d log total_count = gmu mu + gth theta, d / d logits = gmu mu
      const float up = a.upstream, gm = cg.dmu * mu;
      row[g] = up * (gm + cg.dth * th);
      row[G + g] = up * gm;
      if (ZI) row[2 * G + g] = up * cg.dpi;
    } else if (a.train) {
      float up = a.upstream;
      row[G + g] = up * cg.dth * dth_db;
      if (ZI) row[2 * G + g] = up * cg.dpi;
      float gm = up * cg.dmu * dmu_da;         // d loss / d (pre-activation of the mean)
      if (a.scvi) {
        // m = eL * clamp(s_raw): t_g = d loss / d s_raw_g; softmax jacobian finished after the row sum
        float tg = gm * eL * gate;
        row[g] = tg;
        dot += s_raw * tg;
        dl += gm * eL * s;                     // d loss / d lib (through eL), all genes
      } else {
        row[g] = gm;
      }
    } else {
      size_t o = (size_t)r * G + g;
      if (a.out_mean) a.out_mean[o] = mu;
      if (a.out_disp) a.out_disp[o] = th;
      if (ZI && a.out_pi) a.out_pi[o] = pi;
    }
  }
  llk = block_sum(llk, scratch);
  if (threadIdx.x == 0) a.llk_x[r] = llk;
  if (a.train && a.scvi) {
    dot = block_sum(dot, scratch);
    dl = block_sum(dl, scratch);
    // du_g = s_raw_g * (t_g - sum_j s_raw_j t_j); each thread revisits exactly the genes it wrote
    for (int g = threadIdx.x; g < G; g += blockDim.x) row[g] = s_cache[g] * (row[g] - dot);
    if (threadIdx.x == 0 && a.dlib) a.dlib[r] = lib_open ? dl : 0.f;
  }
}

}  // namespace sisua
