// The 64-wide "middle" of the network: BN statistics, BN/bias + ReLU applied on load, the small
// dense layers (hidden x hidden, latent / library / protein heads, first decoder layer), the
// reparameterisation + analytic KL, and their backward passes.  These kernels are shared by the
// un-fused cross-check path and the fused tcgen05 path (which only replaces the genes x hidden
// contractions).  Reference semantics: SURVEY.md section 8a rows a4-a6, a10-a12, Appendix A.
#pragma once
#include "device_math.cuh"

namespace sisua {

constexpr int kH = 64;          // hidden width (single_cell_model.py:78-81)
constexpr int kTileR = 64;      // rows per CTA tile in the mid kernels
constexpr int kMidThreads = 256;

// Pointers that arrive inside by-value argument structs are compiled as generic addresses (LD/ST + address-space
// checks around every atomic); telling the compiler they are global shrinks the code of the once-through kernels.
#define SISUA_GLOBAL(p) __builtin_assume((p) == nullptr || __isGlobal(p))

enum NormMode : int { NORM_RAW = 0, NORM_BN_BATCH = 1, NORM_BN_MOVING = 2, NORM_BIAS = 3 };

struct NormSpec {
  int mode;
  const double* sum;     // [64] column sums of the pre-activation (NORM_BN_BATCH)
  const double* sumsq;   // [64]
  const float* moving;   // [2,64] moving mean, moving var (NORM_BN_MOVING)
  const float* gamma;    // [64]
  const float* beta;     // [64]  (bias when NORM_BIAS)
  float inv_count;       // 1 / rows that entered the batch statistics
  float eps;
  DropSpec drop;         // dropout applied after the ReLU (training only; rate 0 = off)
};

// per-column affine (h = relu(a*sc + sh)) and the standardisation (xhat = (a - mean)*rstd)
// out of line: double-precision divide / sqrt expand to long instruction sequences, and the small kernels that call
// this once per CTA are bound by instruction fetch.  Everything by value: handing out references to a kernel's
// by-value argument struct would force the whole struct into local memory.
__device__ __noinline__ float4 norm_coeffs4(NormSpec ns, int c) {     // (sc, sh, mean, rstd)
  float sc, sh, mean, rstd;
  if (ns.mode == NORM_BN_BATCH) {
    double m = ns.sum[c] * (double)ns.inv_count;
    double v = ns.sumsq[c] * (double)ns.inv_count - m * m;
    if (v < 0.0) v = 0.0;
    mean = (float)m;
    rstd = (float)(1.0 / sqrt(v + (double)ns.eps));
    sc = ns.gamma[c] * rstd;
    sh = ns.beta[c] - mean * sc;
  } else if (ns.mode == NORM_BN_MOVING) {
    mean = ns.moving[c];
    rstd = rsqrtf(ns.moving[kH + c] + ns.eps);
    sc = ns.gamma[c] * rstd;
    sh = ns.beta[c] - mean * sc;
  } else if (ns.mode == NORM_BIAS) {
    mean = 0.f; rstd = 1.f; sc = 1.f; sh = ns.beta[c];
  } else {
    mean = 0.f; rstd = 1.f; sc = 1.f; sh = 0.f;
  }
  return make_float4(sc, sh, mean, rstd);
}
__device__ __forceinline__ void norm_coeffs(const NormSpec& ns, int c, float& sc, float& sh, float& mean, float& rstd) {
  float4 q = norm_coeffs4(ns, c);
  sc = q.x; sh = q.y; mean = q.z; rstd = q.w;
}

// ------------------------------------------------------------------------------------------
// column sums / sums of squares of A[R, ncols<=128] in double (BN batch statistics)
// ------------------------------------------------------------------------------------------
// 16 threads per row (one float4 each), 16 rows per block pass, four row passes in flight; per-thread partial sums in
// fp32 (a handful of rows), combined in fp64.  ncols must be 64 (lda a multiple of 4, A 16-byte aligned).
__global__ void __launch_bounds__(256) col_stats_kernel(const float* __restrict__ A, int lda, int R, int ncols,
                                                        double* __restrict__ sum, double* __restrict__ sumsq) {
  __shared__ float s1[16][kH], s2[16][kH];
  // (no early trigger: co-resident waiting CTAs slowed the running kernel down)
  pdl_wait();
  const int c4 = (threadIdx.x & 15) * 4, rg = threadIdx.x >> 4;
  float a1[4] = {0.f, 0.f, 0.f, 0.f}, a2[4] = {0.f, 0.f, 0.f, 0.f};
  const int stride = gridDim.x * 16;
  auto acc = [&](const float4& v) {
    a1[0] += v.x; a1[1] += v.y; a1[2] += v.z; a1[3] += v.w;
    a2[0] = fmaf(v.x, v.x, a2[0]); a2[1] = fmaf(v.y, v.y, a2[1]); a2[2] = fmaf(v.z, v.z, a2[2]); a2[3] = fmaf(v.w, v.w, a2[3]);
  };
  int r = blockIdx.x * 16 + rg;
  for (; r + 3 * stride < R; r += 4 * stride) {
    const float4 v0 = *reinterpret_cast<const float4*>(A + (size_t)r * lda + c4);
    const float4 v1 = *reinterpret_cast<const float4*>(A + (size_t)(r + stride) * lda + c4);
    const float4 v2 = *reinterpret_cast<const float4*>(A + (size_t)(r + 2 * stride) * lda + c4);
    const float4 v3 = *reinterpret_cast<const float4*>(A + (size_t)(r + 3 * stride) * lda + c4);
    acc(v0); acc(v1); acc(v2); acc(v3);
  }
  for (; r < R; r += stride) acc(*reinterpret_cast<const float4*>(A + (size_t)r * lda + c4));
#pragma unroll
  for (int j = 0; j < 4; ++j) { s1[rg][c4 + j] = a1[j]; s2[rg][c4 + j] = a2[j]; }
  __syncthreads();
  if (threadIdx.x < kH) {
    const int c = threadIdx.x;
    double t1 = 0.0, t2 = 0.0;
    for (int g = 0; g < 16; ++g) { t1 += (double)s1[g][c]; t2 += (double)s2[g][c]; }
    atomicAdd(&sum[c], t1);
    atomicAdd(&sumsq[c], t2);
  }
}

// moving <- m * moving + (1-m) * batch   (one block per BN layer: the tail blocks of elbo_kernel)
struct MovingUpdateArgs {
  const double* sum[16];
  const double* sumsq[16];
  float inv_count[16];
};
// D = relu(norm(A))  (materialises the activated decoder output for the output heads)
__global__ void __launch_bounds__(256) norm_relu_kernel(const float* __restrict__ A, int lda, NormSpec ns,
                                                        float* __restrict__ D, int R) {
  __shared__ float sc[kH], sh[kH];
  // (no early trigger: co-resident waiting CTAs slowed the running kernel down)
  pdl_wait();
  if (threadIdx.x < kH) { float m, r; norm_coeffs(ns, threadIdx.x, sc[threadIdx.x], sh[threadIdx.x], m, r); }
  __syncthreads();
  size_t n = (size_t)R * kH;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    int c = (int)(i % kH); size_t r = i / kH;
    float v = A[r * lda + c] * sc[c] + sh[c];
    if (ns.mode != NORM_RAW) v = fmaxf(v, 0.f) * dropout_mult(ns.drop, (uint32_t)r, (uint32_t)c);
    D[i] = v;
  }
}

// ------------------------------------------------------------------------------------------
// A_out[R, Nout] = act_in(A_in)[R, Kin] . W[Nout, Kin]^T (+ bias);  Kin, Nout <= 64.
// 64-row tiles, 4x4 register blocking (operands k-major in shared memory, two LDS.128 per 16 FMA).
// Optionally accumulates the column sums / sums of squares of A_out (next layer's BN statistics).
// ------------------------------------------------------------------------------------------
constexpr int kTS = 68;   // padded tile stride in floats (16-byte aligned rows)
constexpr size_t kDenseFwdSmem = (size_t)(2 * kH * kTS + 2 * kH + 2 * 16 * kH) * sizeof(float) + 2 * kH * sizeof(double);

__global__ void __launch_bounds__(kMidThreads) dense_fwd_kernel(
    const float* __restrict__ A_in, int lda, int Kin, NormSpec ns, const float* __restrict__ W, int ldw,
    const float* __restrict__ bias, int Nout, float* __restrict__ A_out, int ldo, int R, double* __restrict__ out_sum,
    double* __restrict__ out_sumsq) {
  extern __shared__ __align__(16) float dsm[];
  float* WT = dsm;                       // [k][n]
  float* HT = WT + kH * kTS;             // [k][r]
  float* sc = HT + kH * kTS;
  float* sh = sc + kH;
  float* red1 = sh + kH;                 // [16][64] per-thread-row partial column sums
  float* red2 = red1 + 16 * kH;
  double* dacc = reinterpret_cast<double*>(red2 + 16 * kH);   // [2][64] running sums of this CTA
  const int t = threadIdx.x, ty = t >> 4, tx = t & 15;
  // (no early trigger: co-resident waiting CTAs slowed the running kernel down)
  {   // all 16 loads of a thread are issued before any is used (the kernels are latency-, not bandwidth-bound)
    float w[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      int i = t + kMidThreads * j, n = i >> 6, k = i & 63;
      w[j] = (n < Nout && k < Kin) ? __ldg(W + (size_t)n * ldw + k) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      int i = t + kMidThreads * j, n = i >> 6, k = i & 63;
      WT[k * kTS + n] = w[j];
    }
  }
  pdl_wait();   // weights are step-constant; everything below reads what earlier kernels of this step wrote
  if (t < kH) {
    float m, r;
    if (t < Kin) norm_coeffs(ns, t, sc[t], sh[t], m, r); else { sc[t] = 0.f; sh[t] = 0.f; }
    dacc[t] = 0.0; dacc[kH + t] = 0.0;
  }
  __syncthreads();
  for (int r0 = blockIdx.x * kTileR; r0 < R; r0 += gridDim.x * kTileR) {
    {
      float hv[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        int i = t + kMidThreads * j, r = i >> 6, k = i & 63;
        hv[j] = (r0 + r < R && k < Kin) ? A_in[(size_t)(r0 + r) * lda + k] : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        int i = t + kMidThreads * j, r = i >> 6, k = i & 63;
        float v = 0.f;
        if (r0 + r < R && k < Kin) {
          v = hv[j] * sc[k] + sh[k];
          if (ns.mode != NORM_RAW) v = fmaxf(v, 0.f) * dropout_mult(ns.drop, (uint32_t)(r0 + r), (uint32_t)k);
        }
        HT[k * kTS + r] = v;
      }
    }
    __syncthreads();
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int k = 0; k < Kin; ++k) {
      float4 a4 = *reinterpret_cast<const float4*>(HT + k * kTS + 4 * ty);
      float4 b4 = *reinterpret_cast<const float4*>(WT + k * kTS + 4 * tx);
      float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    float cs[4] = {0.f, 0.f, 0.f, 0.f}, cq[4] = {0.f, 0.f, 0.f, 0.f};
    const bool vec_out = Nout == kH && (ldo & 3) == 0 && (reinterpret_cast<uintptr_t>(A_out) & 15) == 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = r0 + 4 * ty + i;
      float vv[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = 4 * tx + j;
        vv[j] = acc[i][j] + ((bias && n < Nout) ? bias[n] : 0.f);
        if (r < R && n < Nout) {
          if (!vec_out) A_out[(size_t)r * ldo + n] = vv[j];
          cs[j] += vv[j]; cq[j] += vv[j] * vv[j];
        }
      }
      if (vec_out && r < R) *reinterpret_cast<float4*>(A_out + (size_t)r * ldo + 4 * tx) = make_float4(vv[0], vv[1], vv[2], vv[3]);
    }
    if (out_sum) {
#pragma unroll
      for (int j = 0; j < 4; ++j) { red1[ty * kH + 4 * tx + j] = cs[j]; red2[ty * kH + 4 * tx + j] = cq[j]; }
      __syncthreads();
      if (t < kH) {
        double s1 = 0.0, s2 = 0.0;
        for (int y = 0; y < 16; ++y) { s1 += (double)red1[y * kH + t]; s2 += (double)red2[y * kH + t]; }
        dacc[t] += s1; dacc[kH + t] += s2;
      }
    }
    __syncthreads();
  }
  if (out_sum && t < Nout) { atomicAdd(&out_sum[t], dacc[t]); atomicAdd(&out_sumsq[t], dacc[kH + t]); }
}

// ------------------------------------------------------------------------------------------
// latent head: loc/scale, reparameterised sample, analytic KL (rows a5, a6, a10, a14)
// ------------------------------------------------------------------------------------------
struct LatentArgs {
  const float* PL;     // [B, ZP] raw latent projection (ZP = 2Z, or Z for the deterministic DCA latent)
  const float* eps_z;  // [S, B, Z]; null -> Philox normals (noise)
  NoiseSpec noise;
  float* loc;          // [B, Z]
  float* scale;        // [B, Z]
  float* z;            // [S*B, Z]
  float* kl_z;         // [B]
  // scVI library latent
  const float* PLIB;   // [B, 2] raw (loc, scale_raw) or null
  const float* eps_l;  // [S, B]
  const float* library;  // [B, 2] prior (mean, var)
  float* lib_loc; float* lib_scale;  // [B]
  float* lib;          // [S*B] sampled log-library
  float* kl_l;         // [B]
  float* logw;         // [S*B] log p(z_s) - log q(z_s | x) (+ the library latent's), nullable: importance weights of marginal_log_prob
  int B, S, Z, deterministic, scale_act;
};
__global__ void latent_fwd_kernel(LatentArgs a) {
  // (no early trigger: co-resident waiting CTAs slowed the running kernel down)
  pdl_wait();
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.B) return;
  int Z = a.Z;
  if (a.logw && a.PL && a.deterministic)       // (PL == null: the fused latent block already wrote the z part)
    for (int s = 0; s < a.S; ++s) a.logw[(size_t)s * a.B + b] = 0.f;
  if (!a.PL) {
    // library-only call (the z path ran in latent_block_fwd_kernel)
  } else if (a.deterministic) {
    for (int j = 0; j < Z; ++j) {
      float v = a.PL[(size_t)b * Z + j];
      if (a.deterministic == 1) v = fmaxf(v, 0.f);       // 1: 'relu' latent, 2: 'linear' latent (dca.py:16-27)
      a.loc[(size_t)b * Z + j] = v; a.scale[(size_t)b * Z + j] = 0.f;
      for (int s = 0; s < a.S; ++s) a.z[((size_t)s * a.B + b) * Z + j] = v;      // every "sample" of a deterministic latent is the same
    }
    for (int s = 0; s < a.S; ++s) a.kl_z[(size_t)s * a.B + b] = 0.f;
  } else {
    float kl = 0.f;
    if (a.logw) for (int s = 0; s < a.S; ++s) a.logw[(size_t)s * a.B + b] = 0.f;
    for (int j = 0; j < Z; ++j) {
      float mu = a.PL[(size_t)b * 2 * Z + j];
      float sg, dsg;
      activation(a.scale_act, a.PL[(size_t)b * 2 * Z + Z + j], sg, dsg);
      a.loc[(size_t)b * Z + j] = mu; a.scale[(size_t)b * Z + j] = sg;
      const float lsg = logf(sg);
      kl += sg * sg + mu * mu - 1.f - 2.f * lsg;
      for (int s = 0; s < a.S; ++s) {
        const float e = a.eps_z ? a.eps_z[((size_t)s * a.B + b) * Z + j] : philox_normal(a.noise, (uint32_t)b, (uint32_t)j, kNoiseStreamZ + 2u * s);
        const float zz = fmaf(sg, e, mu);
        a.z[((size_t)s * a.B + b) * Z + j] = zz;
        // log N(z; 0, 1) - log N(z; mu, sg) = -z^2/2 + eps^2/2 + log sg
        if (a.logw) a.logw[(size_t)s * a.B + b] += 0.5f * (e * e - zz * zz) + lsg;
      }
    }
    for (int s = 0; s < a.S; ++s) a.kl_z[(size_t)s * a.B + b] = 0.5f * kl;      // same value for every Monte-Carlo sample row
  }
  if (a.PLIB) {
    float mu = a.PLIB[(size_t)b * 2];
    float sg, dsg;
    activation(a.scale_act, a.PLIB[(size_t)b * 2 + 1], sg, dsg);
    float pm = a.library[(size_t)b * 2], pv = a.library[(size_t)b * 2 + 1];
    a.lib_loc[b] = mu; a.lib_scale[b] = sg;
    const float kll = logf(sqrtf(pv) / sg) + (sg * sg + (mu - pm) * (mu - pm)) / (2.f * pv) - 0.5f;
    for (int s = 0; s < a.S; ++s) a.kl_l[(size_t)s * a.B + b] = kll;
    for (int s = 0; s < a.S; ++s) {
      const float e = a.eps_l ? a.eps_l[(size_t)s * a.B + b] : philox_normal(a.noise, (uint32_t)b, 0u, kNoiseStreamL + 2u * s);
      const float l = fmaf(sg, e, mu);
      a.lib[(size_t)s * a.B + b] = l;
      // log N(l; pm, pv) - log N(l; mu, sg)
      if (a.logw) a.logw[(size_t)s * a.B + b] += -0.5f * (l - pm) * (l - pm) / pv - 0.5f * logf(pv) + 0.5f * e * e + logf(sg);
    }
  } else if (a.kl_l) {
    for (int s = 0; s < a.S; ++s) a.kl_l[(size_t)s * a.B + b] = 0.f;
  }
}

// backward of the scVI library latent (row a10): d loss / d sampled log-library -> d(raw loc, raw scale), KL included.
// (The z latent's backward lives in latent_block_bwd_kernel.)
struct LibraryBwdArgs {
  const float* dLib;   // [B] d loss / d sampled log-library
  const float* PLIB; const float* eps_l; const float* library; const float* lib_loc; const float* lib_scale;
  float* dPLIB;        // [B, 2]
  NoiseSpec noise;     // eps_l == null
  int B, scale_act;
  float kl_weight;     // beta / B : d loss / d KL_b
};
__global__ void library_bwd_kernel(LibraryBwdArgs a) {
  pdl_wait();
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.B) return;
  float dl = a.dLib[b];
  float mu = a.lib_loc[b], sg = a.lib_scale[b];
  float pm = a.library[(size_t)b * 2], pv = a.library[(size_t)b * 2 + 1];
  float v, dv;
  activation(a.scale_act, a.PLIB[(size_t)b * 2 + 1], v, dv);
  a.dPLIB[(size_t)b * 2] = dl + a.kl_weight * (mu - pm) / pv;
  const float e = a.eps_l ? a.eps_l[b] : philox_normal(a.noise, (uint32_t)b, 0u, kNoiseStreamL);
  a.dPLIB[(size_t)b * 2 + 1] = (dl * e + a.kl_weight * (sg / pv - 1.f / sg)) * dv;
}

// ------------------------------------------------------------------------------------------
// semi-supervised protein head (row a11): llk_y per row and d loss / d raw head outputs
// ------------------------------------------------------------------------------------------
struct YHeadArgs {
  const float* PY;       // [R, 2P] raw (a | b)
  const float* y;        // [B, P]
  const uint8_t* mask;   // [B] or null (== unlabelled)
  float* llk_y;          // [R]
  float* dPY;            // [R, 2P] or null (inference)
  float* y_mean;         // [R, P] or null
  int R, B, P, y_dist, mean_act, disp_act;
  float upstream;        // -alpha / B   (d loss / d llk_y for a labelled cell, before mask weighting)
  const float* mask_scale;  // device scalar: 1 (Q3 default) or B / n_labelled
};
__global__ void yhead_kernel(YHeadArgs a) {
  // (no early trigger: co-resident waiting CTAs slowed the running kernel down)
  pdl_wait();
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= a.R) return;
  int b = r % a.B, P = a.P;
  float w = (a.mask && a.mask[b]) ? 1.f : 0.f;
  if (a.mask_scale) w *= *a.mask_scale;
  float acc = 0.f;
  for (int j = 0; j < P; ++j) {
    float ra = a.PY[(size_t)r * 2 * P + j], rb = a.PY[(size_t)r * 2 * P + P + j];
    float yv = a.y ? a.y[(size_t)b * P + j] : 0.f;      // (no targets: decode only, the likelihood is discarded)
    float da = 0.f, db = 0.f, llk, mean;
    if (a.y_dist == 0) {
      llk = a.dPY ? nb_tfp_llk<true>(yv, ra, rb, da, db) : nb_tfp_llk<false>(yv, ra, rb, da, db);
      mean = __expf(ra + rb);
    } else {
      float mu, dmu, th, dth;
      activation(a.mean_act, ra, mu, dmu);
      activation(a.disp_act, rb, th, dth);
      CountGrad g;
      llk = count_llk<false, true>(yv, mu, th, 0.f, g);
      da = g.dmu * dmu; db = g.dth * dth; mean = mu;
    }
    acc += llk;
    if (a.dPY) {
      a.dPY[(size_t)r * 2 * P + j] = a.upstream * w * da;
      a.dPY[(size_t)r * 2 * P + P + j] = a.upstream * w * db;
    }
    if (a.y_mean) a.y_mean[(size_t)r * P + j] = mean;
  }
  a.llk_y[r] = acc;
}

// marginal_log_prob (sisua/analysis/posterior.py:941-976 -> odin-ai): per cell
//   mllk_b = logsumexp_s( llk_x[s,b] + alpha * mask_b * llk_y[s,b] + logw[s,b] ) - log S      (importance-weighted bound)
//   llk_x_b = logsumexp_s llk_x[s,b] - log S, same for llk_y                                   (the per-output entries)
struct MarginalArgs {
  const float* terms;    // [5, S*B]
  const float* logw;     // [S*B]
  const uint8_t* mask;   // [B] or null
  float alpha;
  int B, S, has_y;
  float* mllk; float* llk_x; float* llk_y;   // [B] each (llk_y nullable)
};
__global__ void marginal_kernel(MarginalArgs a) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.B) return;
  const size_t R = (size_t)a.S * a.B;
  const float w = (a.has_y && a.mask && a.mask[b]) ? a.alpha : 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY;
  for (int s = 0; s < a.S; ++s) {
    const size_t r = (size_t)s * a.B + b;
    const float lx = a.terms[R + r], ly = a.terms[2 * R + r];
    m0 = fmaxf(m0, lx + w * ly + a.logw[r]); m1 = fmaxf(m1, lx); m2 = fmaxf(m2, ly);
  }
  float s0 = 0.f, s1 = 0.f, s2 = 0.f;
  for (int s = 0; s < a.S; ++s) {
    const size_t r = (size_t)s * a.B + b;
    const float lx = a.terms[R + r], ly = a.terms[2 * R + r];
    s0 += expf(lx + w * ly + a.logw[r] - m0); s1 += expf(lx - m1); s2 += expf(ly - m2);
  }
  const float lS = logf((float)a.S);
  a.mllk[b] = m0 + logf(s0) - lS;
  a.llk_x[b] = m1 + logf(s1) - lS;
  if (a.llk_y) a.llk_y[b] = m2 + logf(s2) - lS;
}

// n_labelled -> mask_scale = B / max(n_labelled, 1)   (Q3 alternative)
__global__ void mask_scale_kernel(const uint8_t* mask, int B, float* out) {
  __shared__ float scratch[33];
  float c = 0.f;
  for (int i = threadIdx.x; i < B; i += blockDim.x) c += mask[i] ? 1.f : 0.f;
  c = block_sum(c, scratch);
  if (threadIdx.x == 0) *out = (float)B / fmaxf(c, 1.f);
}

// ------------------------------------------------------------------------------------------
// ELBO assembly (row a12): terms = [elbo | llk_x | llk_y | kl_z | kl_l] each [R]; loss = -mean
// ------------------------------------------------------------------------------------------
struct ElboArgs {
  float* terms; const uint8_t* mask; const float* mask_scale;
  int R, B; float alpha, beta; float* loss;  // loss: device scalar, pre-zeroed
  // training: the last n_bn blocks of the grid fold the batch statistics into the moving ones (one block per BN layer)
  int n_bn; float* moving; float momentum; MovingUpdateArgs mu;
  // training: sticky word in mapped HOST memory, set when a partial sum of the loss is not finite (SURVEY.md section 5:
  // the reference's terminate_on_nan callback looks at every step; the host reads the word without a synchronisation)
  volatile int* nonfinite;
};
__global__ void __launch_bounds__(256) elbo_kernel(ElboArgs a) {
  // (no early trigger: co-resident waiting CTAs slowed the running kernel down)
  pdl_wait();
  __shared__ float scratch[33];
  const int nblk = gridDim.x - a.n_bn;
  if ((int)blockIdx.x >= nblk) {
    const int l = blockIdx.x - nblk, c = threadIdx.x;
    if (c < kH) {
      double m = a.mu.sum[l][c] * (double)a.mu.inv_count[l];
      double v = a.mu.sumsq[l][c] * (double)a.mu.inv_count[l] - m * m;
      if (v < 0.0) v = 0.0;
      float* mv = a.moving + (size_t)l * 2 * kH;
      mv[c] = a.momentum * mv[c] + (1.f - a.momentum) * (float)m;
      mv[kH + c] = a.momentum * mv[kH + c] + (1.f - a.momentum) * (float)v;
    }
    return;
  }
  float local = 0.f;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < a.R; r += nblk * blockDim.x) {
    int b = r % a.B;
    float w = (a.mask && a.mask[b]) ? 1.f : 0.f;
    if (a.mask_scale) w *= *a.mask_scale;
    float e = a.terms[(size_t)1 * a.R + r] + a.alpha * w * a.terms[(size_t)2 * a.R + r] -
              a.beta * (a.terms[(size_t)3 * a.R + b] + a.terms[(size_t)4 * a.R + b]);
    a.terms[r] = e;
    local += e;
  }
  local = block_sum(local, scratch);
  if (threadIdx.x == 0 && a.loss) atomicAdd(a.loss, -local / (float)a.R);
  if (threadIdx.x == 0 && a.nonfinite && !isfinite(local)) *a.nonfinite = 1;
}

// ------------------------------------------------------------------------------------------
// backward of norm+relu: column sums of dY and dY*xhat (double) + d gamma / d beta
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(
    const float* __restrict__ dH, int ldd, const float* __restrict__ A, int lda, NormSpec ns, int R,
    double* __restrict__ sdy, double* __restrict__ sdyx, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  __shared__ float s1[16][kH], s2[16][kH];
  __shared__ float sc[kH], sh[kH], mean[kH], rstd[kH];
  // (no early trigger: co-resident waiting CTAs slowed the running kernel down)
  pdl_wait();
  if (threadIdx.x < kH) {
    const float4 q = norm_coeffs4(ns, threadIdx.x);
    sc[threadIdx.x] = q.x; sh[threadIdx.x] = q.y; mean[threadIdx.x] = q.z; rstd[threadIdx.x] = q.w;
  }
  __syncthreads();
  // 16 threads per row (one float4 each), 16 rows per block pass, two row passes in flight; fp32 partials per thread
  const int c4 = (threadIdx.x & 15) * 4, rg = threadIdx.x >> 4;
  float a1[4] = {0.f, 0.f, 0.f, 0.f}, a2[4] = {0.f, 0.f, 0.f, 0.f};
  const int stride = gridDim.x * 16;
  const bool drop = ns.drop.rate > 0.f;
  auto acc = [&](const float4& g4, const float4& v4, int r) {
    const float gv[4] = {g4.x, g4.y, g4.z, g4.w}, av[4] = {v4.x, v4.y, v4.z, v4.w};
    float dm[4] = {1.f, 1.f, 1.f, 1.f};
    if (drop) {
      DropMult8 m = dropout_mult8(ns.drop, (uint32_t)r, (uint32_t)(c4 >> 3));
#pragma unroll
      for (int j = 0; j < 4; ++j) dm[j] = m.m[(c4 & 4) + j];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = c4 + j;
      const float dy = (av[j] * sc[c] + sh[c] > 0.f) ? gv[j] * dm[j] : 0.f;
      a1[j] += dy;
      a2[j] = fmaf(dy, (av[j] - mean[c]) * rstd[c], a2[j]);
    }
  };
  int r = blockIdx.x * 16 + rg;
  for (; r + stride < R; r += 2 * stride) {
    const float4 g0 = *reinterpret_cast<const float4*>(dH + (size_t)r * ldd + c4);
    const float4 v0 = *reinterpret_cast<const float4*>(A + (size_t)r * lda + c4);
    const float4 g1 = *reinterpret_cast<const float4*>(dH + (size_t)(r + stride) * ldd + c4);
    const float4 v1 = *reinterpret_cast<const float4*>(A + (size_t)(r + stride) * lda + c4);
    acc(g0, v0, r); acc(g1, v1, r + stride);
  }
  for (; r < R; r += stride)
    acc(*reinterpret_cast<const float4*>(dH + (size_t)r * ldd + c4), *reinterpret_cast<const float4*>(A + (size_t)r * lda + c4), r);
#pragma unroll
  for (int j = 0; j < 4; ++j) { s1[rg][c4 + j] = a1[j]; s2[rg][c4 + j] = a2[j]; }
  __syncthreads();
  if (threadIdx.x < kH) {
    const int c = threadIdx.x;
    double t1 = 0.0, t2 = 0.0;
    for (int g = 0; g < 16; ++g) { t1 += (double)s1[g][c]; t2 += (double)s2[g][c]; }
    atomicAdd(&sdy[c], t1);
    atomicAdd(&sdyx[c], t2);
    if (ns.mode == NORM_BIAS) {
      atomicAdd(&dbeta[c], (float)t1);             // bias gradient
    } else {
      atomicAdd(&dgamma[c], (float)t2);
      atomicAdd(&dbeta[c], (float)t1);
    }
  }
}

// ------------------------------------------------------------------------------------------
// backward of one dense unit:  given the gradient at its output, produce dW (atomic), db (atomic),
// the gradient wrt its input activations, and optionally the gradient wrt its pre-activation.
//   out_mode 0: dOut[R, Nout] is given directly (heads with bias, no norm)
//   out_mode 1: dOut = BN/bias backward of dH[R,64] applied on load (needs A_out, ns_out, sdy, sdyx)
// When `prev_*` is set the kernel also reduces, for the unit feeding this one, the column sums of
// dY and dY*xhat that ITS BatchNorm backward needs (plus its d gamma / d beta), so no separate pass
// over dIn is required.  64-row tiles, 4x4 register blocking.
// ------------------------------------------------------------------------------------------
constexpr size_t kDenseBwdSmem =
    (size_t)(5 * kH * kTS + 11 * kH + 2 * 16 * kH) * sizeof(float) + 2 * kH * sizeof(double);
struct DenseBwdArgs {
  int out_mode;
  const float* dOut; int ldd;          // dOut (mode 0) or dH (mode 1)
  const float* A_out; int lda_out;     // pre-activation of this unit (mode 1)
  NormSpec ns_out;
  const double* sdy; const double* sdyx;
  int Nout;
  const float* A_in; int lda_in; int Kin; NormSpec ns_in;   // source of the unit's input (null -> no dW)
  const float* W; int ldw;             // [Nout, Kin]
  float* dW;                           // [Nout, Kin] ld = ldw, atomic accumulate (nullable)
  float* db;                           // [Nout] atomic accumulate (nullable)
  float* dIn; int ldi; int accumulate_dIn;   // [R, Kin] (nullable)
  float* dA; int ldda;                 // [R, Nout] d loss / d pre-activation (nullable)
  int R;
  // fused reduction for the producing unit's norm backward (nullable)
  double* prev_sdy; double* prev_sdyx; float* prev_dgamma; float* prev_dbeta;
};
__global__ void __launch_bounds__(kMidThreads) dense_bwd_kernel(DenseBwdArgs a) {
  SISUA_GLOBAL(a.dOut); SISUA_GLOBAL(a.A_out); SISUA_GLOBAL(a.sdy); SISUA_GLOBAL(a.sdyx); SISUA_GLOBAL(a.A_in); SISUA_GLOBAL(a.W);
  SISUA_GLOBAL(a.dW); SISUA_GLOBAL(a.db); SISUA_GLOBAL(a.dIn); SISUA_GLOBAL(a.dA); SISUA_GLOBAL(a.prev_sdy); SISUA_GLOBAL(a.prev_sdyx);
  SISUA_GLOBAL(a.prev_dgamma); SISUA_GLOBAL(a.prev_dbeta);
  SISUA_GLOBAL(a.ns_out.sum); SISUA_GLOBAL(a.ns_out.sumsq); SISUA_GLOBAL(a.ns_out.gamma); SISUA_GLOBAL(a.ns_out.beta); SISUA_GLOBAL(a.ns_out.moving);
  SISUA_GLOBAL(a.ns_in.sum); SISUA_GLOBAL(a.ns_in.sumsq); SISUA_GLOBAL(a.ns_in.gamma); SISUA_GLOBAL(a.ns_in.beta); SISUA_GLOBAL(a.ns_in.moving);
  extern __shared__ __align__(16) float dyn_smem[];
  float* Wn = dyn_smem;                 // [n][k]
  float* Hn = Wn + kH * kTS;            // [r][k]   activated input
  float* Gn = Hn + kH * kTS;            // [r][n]   gradient wrt this unit's pre-activation
  float* GT = Gn + kH * kTS;            // [n][r]
  float* An = GT + kH * kTS;            // [r][k]   raw pre-activation of the producing unit (fused reduction)
  float* sc_i = An + kH * kTS;
  float *sh_i = sc_i + kH, *mean_i = sh_i + kH, *rstd_i = mean_i + kH, *sc_o = rstd_i + kH, *sh_o = sc_o + kH,
        *mean_o = sh_o + kH, *rstd_o = mean_o + kH, *m1 = rstd_o + kH, *m2 = m1 + kH, *gsc = m2 + kH;
  float* red1 = gsc + kH;
  float* red2 = red1 + 16 * kH;
  double* dacc = reinterpret_cast<double*>(red2 + 16 * kH);
  const int t = threadIdx.x, ty = t >> 4, tx = t & 15;
  const int Nout = a.Nout, Kin = a.Kin;
  const bool has_in = a.A_in != nullptr;
  const bool fuse_prev = a.prev_sdy != nullptr;
  // (no early trigger: co-resident waiting CTAs slowed the running kernel down)
  {
    float w[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      int i = t + kMidThreads * j, n = i >> 6, k = i & 63;
      w[j] = (n < Nout && k < Kin && a.W) ? __ldg(a.W + (size_t)n * a.ldw + k) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      int i = t + kMidThreads * j, n = i >> 6, k = i & 63;
      Wn[n * kTS + k] = w[j];
    }
  }
  pdl_wait();
  if (t < kH) {
    if (has_in && t < Kin) norm_coeffs(a.ns_in, t, sc_i[t], sh_i[t], mean_i[t], rstd_i[t]);
    else { sc_i[t] = 0.f; sh_i[t] = 0.f; mean_i[t] = 0.f; rstd_i[t] = 0.f; }
    if (a.out_mode == 1) {
      norm_coeffs(a.ns_out, t, sc_o[t], sh_o[t], mean_o[t], rstd_o[t]);
      if (a.ns_out.mode == NORM_BN_BATCH) {
        m1[t] = (float)(a.sdy[t] * (double)a.ns_out.inv_count);
        m2[t] = (float)(a.sdyx[t] * (double)a.ns_out.inv_count);
      } else {                      // bias (or moving-stat BN): no batch coupling
        m1[t] = 0.f; m2[t] = 0.f;
      }
      gsc[t] = sc_o[t];             // gamma * rstd
    }
    dacc[t] = 0.0; dacc[kH + t] = 0.0;
  }
  __syncthreads();
  float accW[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) accW[i][j] = 0.f;
  float accb = 0.f;
  const float in_drop_scale = a.ns_in.drop.rate > 0.f ? a.ns_in.drop.scale : 1.f;
  for (int r0 = blockIdx.x * kTileR; r0 < a.R; r0 += gridDim.x * kTileR) {
    {
      float go[16], ao[16], hi[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        int i = t + kMidThreads * j, r = i >> 6, n = i & 63;
        const bool ok = r0 + r < a.R && n < Nout;
        go[j] = ok ? a.dOut[(size_t)(r0 + r) * a.ldd + n] : 0.f;
        ao[j] = (ok && a.out_mode == 1) ? a.A_out[(size_t)(r0 + r) * a.lda_out + n] : 0.f;
        hi[j] = (has_in && r0 + r < a.R && n < Kin) ? a.A_in[(size_t)(r0 + r) * a.lda_in + n] : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        int i = t + kMidThreads * j, r = i >> 6, n = i & 63;
        float g = 0.f;
        if (r0 + r < a.R && n < Nout) {
          if (a.out_mode == 0) {
            g = go[j];
          } else {
            float av = ao[j];
            float dy = (av * sc_o[n] + sh_o[n] > 0.f) ? go[j] * dropout_mult(a.ns_out.drop, (uint32_t)(r0 + r), (uint32_t)n) : 0.f;
            float xh = (av - mean_o[n]) * rstd_o[n];
            g = gsc[n] * (dy - m1[n] - xh * m2[n]);
          }
          if (a.dA) a.dA[(size_t)(r0 + r) * a.ldda + n] = g;
        }
        Gn[r * kTS + n] = g;
        GT[n * kTS + r] = g;
        if (has_in) {
          const int k = n;
          float v = 0.f, raw = 0.f;
          if (r0 + r < a.R && k < Kin) {
            raw = hi[j];
            v = raw * sc_i[k] + sh_i[k];
            if (a.ns_in.mode != NORM_RAW) v = fmaxf(v, 0.f) * dropout_mult(a.ns_in.drop, (uint32_t)(r0 + r), (uint32_t)k);
          }
          Hn[r * kTS + k] = v;
          if (fuse_prev) An[r * kTS + k] = raw;
        }
      }
    }
    __syncthreads();
    if (has_in && a.dW) {   // dW[n][k] += sum_r G[r][n] * H[r][k];  n = 4ty.., k = 4tx..
      for (int r = 0; r < kTileR; ++r) {
        float4 a4 = *reinterpret_cast<const float4*>(Gn + r * kTS + 4 * ty);
        float4 b4 = *reinterpret_cast<const float4*>(Hn + r * kTS + 4 * tx);
        float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) accW[i][j] = fmaf(av[i], bv[j], accW[i][j]);
      }
    }
    if (a.db && t < Nout) {
      for (int r = 0; r < kTileR; ++r) accb += Gn[r * kTS + t];
    }
    if (a.dIn) {            // dIn[r][k] = sum_n G[r][n] * W[n][k];  r = 4ty.., k = 4tx..
      float acc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
      for (int n = 0; n < Nout; ++n) {
        float4 a4 = *reinterpret_cast<const float4*>(GT + n * kTS + 4 * ty);
        float4 b4 = *reinterpret_cast<const float4*>(Wn + n * kTS + 4 * tx);
        float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      float cs[4] = {0.f, 0.f, 0.f, 0.f}, cq[4] = {0.f, 0.f, 0.f, 0.f};
      const bool vec_in = Kin == kH && (a.ldi & 3) == 0 && (reinterpret_cast<uintptr_t>(a.dIn) & 15) == 0;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int rl = 4 * ty + i, r = r0 + rl;
        if (vec_in && r < a.R) {
          float4* p4 = reinterpret_cast<float4*>(a.dIn + (size_t)r * a.ldi + 4 * tx);
          if (a.accumulate_dIn) {
            const float4 o = *p4;
            acc[i][0] += o.x; acc[i][1] += o.y; acc[i][2] += o.z; acc[i][3] += o.w;
          }
          *p4 = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int k = 4 * tx + j;
          if (r < a.R && k < Kin) {
            float v = acc[i][j];
            if (!vec_in) {
              float* p = a.dIn + (size_t)r * a.ldi + k;
              v = a.accumulate_dIn ? (*p + acc[i][j]) : acc[i][j];
              *p = v;
            }
            if (fuse_prev) {
              // gradient wrt the producing unit's norm output: relu mask (and dropout scale) of h = Hn
              float dy = Hn[rl * kTS + k] > 0.f ? v * in_drop_scale : 0.f;
              float xh = (An[rl * kTS + k] - mean_i[k]) * rstd_i[k];
              cs[j] += dy; cq[j] += dy * xh;
            }
          }
        }
      }
      if (fuse_prev) {
#pragma unroll
        for (int j = 0; j < 4; ++j) { red1[ty * kH + 4 * tx + j] = cs[j]; red2[ty * kH + 4 * tx + j] = cq[j]; }
        __syncthreads();
        if (t < kH) {
          double s1 = 0.0, s2 = 0.0;
          for (int y = 0; y < 16; ++y) { s1 += (double)red1[y * kH + t]; s2 += (double)red2[y * kH + t]; }
          dacc[t] += s1; dacc[kH + t] += s2;
        }
      }
    }
    __syncthreads();
  }
  if (has_in && a.dW) {
    // full 64-wide rows that start on 16-byte boundaries: one vector reduction per four weights
    const bool vec = Kin == kH && (a.ldw & 3) == 0 && (reinterpret_cast<uintptr_t>(a.dW) & 15) == 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int n = 4 * ty + i;
      if (vec) {
        if (n < Nout) red_add_v4f(&a.dW[(size_t)n * a.ldw + 4 * tx], accW[i][0], accW[i][1], accW[i][2], accW[i][3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int k = 4 * tx + j;
          if (n < Nout && k < Kin) atomicAdd(&a.dW[(size_t)n * a.ldw + k], accW[i][j]);
        }
      }
    }
  }
  if (a.db && t < Nout) atomicAdd(&a.db[t], accb);
  if (fuse_prev && t < Kin) {
    atomicAdd(&a.prev_sdy[t], dacc[t]);
    atomicAdd(&a.prev_sdyx[t], dacc[kH + t]);
    if (a.ns_in.mode == NORM_BIAS) {
      atomicAdd(&a.prev_dbeta[t], (float)dacc[t]);
    } else {
      atomicAdd(&a.prev_dgamma[t], (float)dacc[kH + t]);
      atomicAdd(&a.prev_dbeta[t], (float)dacc[t]);
    }
  }
}

// ------------------------------------------------------------------------------------------
// norm + ReLU backward only: dA[r][c] = gamma*rstd * (dy - mean(dy) - xhat * mean(dy*xhat)), dy = dH * relu' * dropout.
// Used for the first encoder unit, whose weight gradient is the big genes x hidden GEMM (enc_first_bwd_kernel).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const float* __restrict__ dH, int ldd, const float* __restrict__ A, int lda,
                                                           NormSpec ns, const double* __restrict__ sdy,
                                                           const double* __restrict__ sdyx, float* __restrict__ dA, int ldda, int R) {
  __shared__ float sc[kH], sh[kH], mean[kH], rstd[kH], m1[kH], m2[kH];
  pdl_wait();
  if (threadIdx.x < kH) {
    const int c = threadIdx.x;
    float4 q = norm_coeffs4(ns, c);
    sc[c] = q.x; sh[c] = q.y; mean[c] = q.z; rstd[c] = q.w;
    const bool bn = ns.mode == NORM_BN_BATCH;
    m1[c] = bn ? (float)(sdy[c] * (double)ns.inv_count) : 0.f;
    m2[c] = bn ? (float)(sdyx[c] * (double)ns.inv_count) : 0.f;
  }
  __syncthreads();
  const int n4 = R * (kH / 4);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
    const int r = i / (kH / 4), c0 = (i % (kH / 4)) * 4;
    const float4 g4 = *reinterpret_cast<const float4*>(dH + (size_t)r * ldd + c0);
    const float4 a4 = *reinterpret_cast<const float4*>(A + (size_t)r * lda + c0);
    const float gv[4] = {g4.x, g4.y, g4.z, g4.w}, av[4] = {a4.x, a4.y, a4.z, a4.w};
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = c0 + j;
      float dy = (av[j] * sc[c] + sh[c] > 0.f) ? gv[j] * dropout_mult(ns.drop, (uint32_t)r, (uint32_t)c) : 0.f;
      o[j] = sc[c] * (dy - m1[c] - (av[j] - mean[c]) * rstd[c] * m2[c]);
    }
    *reinterpret_cast<float4*>(dA + (size_t)r * ldda + c0) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// ------------------------------------------------------------------------------------------
// Fused "latent block": the three row-local steps around the reparameterisation in one kernel.
//   forward : PL = act(A_enc_last) . W_lat^T + b  ->  loc / scale / z = loc + scale*eps / KL  ->  A_dec0 = z . W_dec0^T
//   backward: BN-bwd(dH_dec0) -> dW_dec0, dz -> d(loc, scale) (+ KL gradient) -> dW_lat, db_lat, dH_enc_last
// (no BatchNorm sits between them, so nothing but launch latency separated the former three kernels).
// Same 64-row tiles / 4x4 register blocking as dense_fwd_kernel / dense_bwd_kernel.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void tile_mma(const float* __restrict__ At, const float* __restrict__ Bt, int K, int ty, int tx,
                                         float (&acc)[4][4]) {
  for (int k = 0; k < K; ++k) {
    float4 a4 = *reinterpret_cast<const float4*>(At + k * kTS + 4 * ty);
    float4 b4 = *reinterpret_cast<const float4*>(Bt + k * kTS + 4 * tx);
    float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
  }
}
// Narrow variants for the latent block (latent width <= 16, so one side of the product has <= 16 / <= 32 useful
// columns and the 4x4 blocking would spend 3/4 of its FMAs on padding): all 256 threads share the narrow side.
//   tile_mma_n16: C[4 rg + i][c]       = sum_k At[k][4 rg + i] * Bt[k][c],       rg = t >> 4, c = t & 15
//   tile_mma_n32: C[4 rg + i][2 cp + j] = sum_k At[k][4 rg + i] * Bt[k][2 cp + j], rg = t >> 4, cp = t & 15
//   tile_mma_m32: C[4 mg + i][2 kp + j] = sum_k At[k][4 mg + i] * Bt[k][2 kp + j], mg = t >> 5 (rows < 32), kp = t & 31
__device__ __forceinline__ void tile_mma_n16(const float* __restrict__ At, const float* __restrict__ Bt, int bstride, int K, int t,
                                             float (&acc)[4]) {
  const int rg = t >> 4, c = t & 15;
  for (int k = 0; k < K; ++k) {
    const float4 a4 = *reinterpret_cast<const float4*>(At + k * kTS + 4 * rg);
    const float b = Bt[k * bstride + c];
    acc[0] = fmaf(a4.x, b, acc[0]); acc[1] = fmaf(a4.y, b, acc[1]); acc[2] = fmaf(a4.z, b, acc[2]); acc[3] = fmaf(a4.w, b, acc[3]);
  }
}
__device__ __forceinline__ void tile_mma_n32(const float* __restrict__ At, const float* __restrict__ Bt, int K, int t, float (&acc)[4][2]) {
  const int rg = t >> 4, cp = t & 15;
  for (int k = 0; k < K; ++k) {
    const float4 a4 = *reinterpret_cast<const float4*>(At + k * kTS + 4 * rg);
    const float2 b2 = *reinterpret_cast<const float2*>(Bt + k * kTS + 2 * cp);
    const float av[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { acc[i][0] = fmaf(av[i], b2.x, acc[i][0]); acc[i][1] = fmaf(av[i], b2.y, acc[i][1]); }
  }
}
__device__ __forceinline__ void tile_mma_m32(const float* __restrict__ At, const float* __restrict__ Bt, int K, int t, float (&acc)[4][2]) {
  const int mg = t >> 5, kp = t & 31;
  for (int k = 0; k < K; ++k) {
    const float4 a4 = *reinterpret_cast<const float4*>(At + k * kTS + 4 * mg);
    const float2 b2 = *reinterpret_cast<const float2*>(Bt + k * kTS + 2 * kp);
    const float av[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { acc[i][0] = fmaf(av[i], b2.x, acc[i][0]); acc[i][1] = fmaf(av[i], b2.y, acc[i][1]); }
  }
}
__device__ __forceinline__ void zero16(float (&acc)[4][4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
}

struct LatentBlockFwdArgs {
  const float* A_enc; int lda; NormSpec ns_enc;     // last encoder unit (pre-activation + norm)
  const float* W_lat; const float* b_lat; int ZP;   // [ZP, 64]
  const float* eps_z;                               // [B, Z]; null -> Philox normals (noise)
  NoiseSpec noise;
  const float* W_d0;                                // [64, Z]
  float* PL; float* loc; float* scale; float* z; float* kl_z;
  float* logw;                                      // [B] log p(z) - log q(z | x), nullable
  float* A_d0; int ldd0;                            // [B, 64]
  double* out_sum; double* out_sumsq;               // BN statistics of A_d0 (nullable)
  int B, Z, deterministic, scale_act;
};
constexpr size_t kLatentFwdSmem = (size_t)(5 * kH * kTS + 2 * kH + 2 * 16 * kH) * sizeof(float) + 2 * kH * sizeof(double);

__global__ void __launch_bounds__(kMidThreads) latent_block_fwd_kernel(LatentBlockFwdArgs a) {
  SISUA_GLOBAL(a.A_enc); SISUA_GLOBAL(a.W_lat); SISUA_GLOBAL(a.b_lat); SISUA_GLOBAL(a.eps_z); SISUA_GLOBAL(a.W_d0); SISUA_GLOBAL(a.PL);
  SISUA_GLOBAL(a.loc); SISUA_GLOBAL(a.scale); SISUA_GLOBAL(a.z); SISUA_GLOBAL(a.kl_z); SISUA_GLOBAL(a.A_d0); SISUA_GLOBAL(a.out_sum);
  SISUA_GLOBAL(a.out_sumsq);
  extern __shared__ __align__(16) float lsm[];
  float* HT = lsm;                    // [k][r]
  float* W1T = HT + kH * kTS;         // [k][m]   W_lat transposed
  float* PLs = W1T + kH * kTS;        // [r][m]
  float* ZT = PLs + kH * kTS;         // [j][r]
  float* W2T = ZT + kH * kTS;         // [j][n]   W_dec0 transposed
  float* sc = W2T + kH * kTS;
  float* sh = sc + kH;
  float* red1 = sh + kH;
  float* red2 = red1 + 16 * kH;
  double* dacc = reinterpret_cast<double*>(red2 + 16 * kH);
  const int t = threadIdx.x, ty = t >> 4, tx = t & 15;
  const int Z = a.Z, ZP = a.ZP;
  {
    float w1[16], w2[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      int i = t + kMidThreads * j, m = i >> 6, k = i & 63;
      w1[j] = m < ZP ? __ldg(a.W_lat + (size_t)m * kH + k) : 0.f;       // (m, k)
      w2[j] = k < Z ? __ldg(a.W_d0 + (size_t)m * Z + k) : 0.f;          // (n = m, j = k)
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      int i = t + kMidThreads * j, m = i >> 6, k = i & 63;
      W1T[k * kTS + m] = w1[j];
      W2T[k * kTS + m] = w2[j];
    }
  }
  pdl_wait();
  if (t < kH) {
    float m, r;
    norm_coeffs(a.ns_enc, t, sc[t], sh[t], m, r);
    dacc[t] = 0.0; dacc[kH + t] = 0.0;
  }
  __syncthreads();
  for (int r0 = blockIdx.x * kTileR; r0 < a.B; r0 += gridDim.x * kTileR) {
    {
      float hv[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        int i = t + kMidThreads * j, r = i >> 6, k = i & 63;
        hv[j] = r0 + r < a.B ? a.A_enc[(size_t)(r0 + r) * a.lda + k] : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        int i = t + kMidThreads * j, r = i >> 6, k = i & 63;
        float v = 0.f;
        if (r0 + r < a.B) {
          v = hv[j] * sc[k] + sh[k];
          if (a.ns_enc.mode != NORM_RAW) v = fmaxf(v, 0.f) * dropout_mult(a.ns_enc.drop, (uint32_t)(r0 + r), (uint32_t)k);
        }
        HT[k * kTS + r] = v;
      }
    }
    __syncthreads();
    float acc[4][4];
    if (ZP <= 32) {          // only the first 32 columns of PL are meaningful: 8 outputs per thread instead of 16
      float pn[4][2];
#pragma unroll
      for (int i = 0; i < 4; ++i) { pn[i][0] = 0.f; pn[i][1] = 0.f; }
      tile_mma_n32(HT, W1T, kH, t, pn);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int r = 4 * ty + i, m = 2 * tx + j;
          float v = m < ZP ? pn[i][j] + a.b_lat[m] : 0.f;
          PLs[r * kTS + m] = v;
          if (r0 + r < a.B && m < ZP) a.PL[(size_t)(r0 + r) * ZP + m] = v;
        }
    } else {
      zero16(acc);
      tile_mma(HT, W1T, kH, ty, tx, acc);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int r = 4 * ty + i, m = 4 * tx + j;
          float v = m < ZP ? acc[i][j] + a.b_lat[m] : 0.f;
          PLs[r * kTS + m] = v;
          if (r0 + r < a.B && m < ZP) a.PL[(size_t)(r0 + r) * ZP + m] = v;
        }
    }
    __syncthreads();
    if (t < kTileR) {      // one thread per row: loc / scale / sample / KL
      const int r = t, b = r0 + r;
      float kl = 0.f, lw = 0.f;
      float4 nz4 = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int j = 0; j < Z; ++j) {
        float mu, sg, zz;
        if (a.deterministic) {
          mu = PLs[r * kTS + j];
          if (a.deterministic == 1) mu = fmaxf(mu, 0.f);
          sg = 0.f; zz = mu;
        } else {
          float dsg;
          mu = PLs[r * kTS + j];
          activation(a.scale_act, PLs[r * kTS + Z + j], sg, dsg);
          float e = 0.f;
          if (b < a.B) {
            if (a.eps_z) {
              e = a.eps_z[(size_t)b * Z + j];
            } else {                      // four normals per Philox call
              if ((j & 3) == 0) nz4 = philox_normal4(a.noise, (uint32_t)b, (uint32_t)(j >> 2), kNoiseStreamZ);
              const int k = j & 3;
              e = k == 0 ? nz4.x : (k == 1 ? nz4.y : (k == 2 ? nz4.z : nz4.w));
            }
          }
          zz = b < a.B ? fmaf(sg, e, mu) : 0.f;
          const float lsg = logf(sg);
          kl += sg * sg + mu * mu - 1.f - 2.f * lsg;
          lw += 0.5f * (e * e - zz * zz) + lsg;
        }
        ZT[j * kTS + r] = b < a.B ? zz : 0.f;
        if (b < a.B) { a.loc[(size_t)b * Z + j] = mu; a.scale[(size_t)b * Z + j] = sg; a.z[(size_t)b * Z + j] = zz; }
      }
      if (b < a.B) a.kl_z[b] = a.deterministic ? 0.f : 0.5f * kl;
      if (b < a.B && a.logw) a.logw[b] = a.deterministic ? 0.f : lw;
    }
    __syncthreads();
    zero16(acc);
    tile_mma(ZT, W2T, Z, ty, tx, acc);
    float cs[4] = {0.f, 0.f, 0.f, 0.f}, cq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = r0 + 4 * ty + i;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (r < a.B) {
          float v = acc[i][j];
          if (a.ldd0 & 3) a.A_d0[(size_t)r * a.ldd0 + 4 * tx + j] = v;
          cs[j] += v; cq[j] += v * v;
        }
      }
      if (r < a.B && (a.ldd0 & 3) == 0)
        *reinterpret_cast<float4*>(a.A_d0 + (size_t)r * a.ldd0 + 4 * tx) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    }
    if (a.out_sum) {
#pragma unroll
      for (int j = 0; j < 4; ++j) { red1[ty * kH + 4 * tx + j] = cs[j]; red2[ty * kH + 4 * tx + j] = cq[j]; }
      __syncthreads();
      if (t < kH) {
        double s1 = 0.0, s2 = 0.0;
        for (int y = 0; y < 16; ++y) { s1 += (double)red1[y * kH + t]; s2 += (double)red2[y * kH + t]; }
        dacc[t] += s1; dacc[kH + t] += s2;
      }
    }
    __syncthreads();
  }
  if (a.out_sum && t < kH) { atomicAdd(&a.out_sum[t], dacc[t]); atomicAdd(&a.out_sumsq[t], dacc[kH + t]); }
}

struct LatentBlockBwdArgs {
  const float* dH_d0;                       // [B, 64] gradient wrt the activated output of decoder unit 0
  const float* A_d0; int ldd0; NormSpec ns_d0; const double* sdy; const double* sdyx;
  const float* W_d0; float* dW_d0;          // [64, Z]
  const float* z; const float* PL; const float* eps_z; const float* loc; const float* scale;
  NoiseSpec noise;                          // eps_z == null
  const float* W_lat; float* dW_lat; float* db_lat; int ZP;     // [ZP, 64]
  const float* A_enc; int lda; NormSpec ns_enc;
  float* dH_enc;                            // [B, 64] gradient wrt the activated output of the last encoder unit
  double* prev_sdy; double* prev_sdyx; float* prev_dgamma; float* prev_dbeta;   // its norm-backward reductions
  int B, Z, deterministic, scale_act;
  float kl_weight;
};
constexpr int kNarrowTS = 20;     // row stride of the two tiles that only hold <= 16 latent columns (narrow path)
constexpr size_t kLatentBwdSmem = (size_t)(7 * kH * kTS + 11 * kH + 2 * 16 * kH) * sizeof(float) + 2 * kH * sizeof(double);
// latent width <= 16: 104 KB instead of 129 KB, so that two CTAs fit on an SM
constexpr size_t kLatentBwdSmemNarrow =
    (size_t)(5 * kH * kTS + 2 * kH * kNarrowTS + 11 * kH + 2 * 16 * kH) * sizeof(float) + 2 * kH * sizeof(double);

__global__ void __launch_bounds__(kMidThreads, 2) latent_block_bwd_kernel(LatentBlockBwdArgs a) {
  SISUA_GLOBAL(a.dH_d0); SISUA_GLOBAL(a.A_d0); SISUA_GLOBAL(a.sdy); SISUA_GLOBAL(a.sdyx); SISUA_GLOBAL(a.W_d0); SISUA_GLOBAL(a.dW_d0);
  SISUA_GLOBAL(a.z); SISUA_GLOBAL(a.PL); SISUA_GLOBAL(a.eps_z); SISUA_GLOBAL(a.loc); SISUA_GLOBAL(a.scale); SISUA_GLOBAL(a.W_lat);
  SISUA_GLOBAL(a.dW_lat); SISUA_GLOBAL(a.db_lat); SISUA_GLOBAL(a.A_enc); SISUA_GLOBAL(a.dH_enc); SISUA_GLOBAL(a.prev_sdy);
  SISUA_GLOBAL(a.prev_sdyx); SISUA_GLOBAL(a.prev_dgamma); SISUA_GLOBAL(a.prev_dbeta);
  extern __shared__ __align__(16) float lsm[];
  float* Gn = lsm;                    // [r][n]  (stage 1: decoder-0 pre-activation gradient; stage 2: dPL)
  float* GT = Gn + kH * kTS;          // [n][r]
  const int zs = a.Z <= 16 ? kNarrowTS : kTS;   // the launch sizes the dynamic shared memory accordingly
  float* Zn = GT + kH * kTS;          // [r][j]  z, later dz
  float* Wd0n = Zn + kH * zs;         // [n][j]
  float* Hn = Wd0n + kH * zs;         // [r][k]  activated encoder output
  float* An = Hn + kH * kTS;          // [r][k]  its raw pre-activation
  float* Wln = An + kH * kTS;         // [m][k]  W_lat
  float* sc_i = Wln + kH * kTS;
  float *sh_i = sc_i + kH, *mean_i = sh_i + kH, *rstd_i = mean_i + kH, *sc_o = rstd_i + kH, *sh_o = sc_o + kH,
        *mean_o = sh_o + kH, *rstd_o = mean_o + kH, *m1 = rstd_o + kH, *m2 = m1 + kH, *gsc = m2 + kH;
  float* red1 = gsc + kH;
  float* red2 = red1 + 16 * kH;
  double* dacc = reinterpret_cast<double*>(red2 + 16 * kH);
  const int t = threadIdx.x, ty = t >> 4, tx = t & 15;
  const int Z = a.Z, ZP = a.ZP;
  {
    float w1[16], w2[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      int i = t + kMidThreads * j, n = i >> 6, k = i & 63;
      w1[j] = k < Z ? __ldg(a.W_d0 + (size_t)n * Z + k) : 0.f;
      w2[j] = n < ZP ? __ldg(a.W_lat + (size_t)n * kH + k) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      int i = t + kMidThreads * j, n = i >> 6, k = i & 63;
      if (k < zs) Wd0n[n * zs + k] = w1[j];
      Wln[n * kTS + k] = w2[j];
    }
  }
  pdl_wait();
  if (t < kH) {
    norm_coeffs(a.ns_enc, t, sc_i[t], sh_i[t], mean_i[t], rstd_i[t]);
    norm_coeffs(a.ns_d0, t, sc_o[t], sh_o[t], mean_o[t], rstd_o[t]);
    if (a.ns_d0.mode == NORM_BN_BATCH) {
      m1[t] = (float)(a.sdy[t] * (double)a.ns_d0.inv_count);
      m2[t] = (float)(a.sdyx[t] * (double)a.ns_d0.inv_count);
    } else { m1[t] = 0.f; m2[t] = 0.f; }
    gsc[t] = sc_o[t];
    dacc[t] = 0.0; dacc[kH + t] = 0.0;
  }
  __syncthreads();
  float accW0[4][4], accW1[4][4];
  zero16(accW0); zero16(accW1);
  // latent width <= 16: the three products with a narrow side use all 256 threads on the useful columns
  const bool narrow = Z <= 16;
  float nW0[4] = {0.f, 0.f, 0.f, 0.f}, nW1[4][2] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
  float accb = 0.f;
  const float in_drop_scale = a.ns_enc.drop.rate > 0.f ? a.ns_enc.drop.scale : 1.f;
  for (int r0 = blockIdx.x * kTileR; r0 < a.B; r0 += gridDim.x * kTileR) {
    {
      float go[16], ao[16], hi[16], zv[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        int i = t + kMidThreads * j, r = i >> 6, n = i & 63;
        const bool ok = r0 + r < a.B;
        go[j] = ok ? a.dH_d0[(size_t)(r0 + r) * kH + n] : 0.f;
        ao[j] = ok ? a.A_d0[(size_t)(r0 + r) * a.ldd0 + n] : 0.f;
        hi[j] = ok ? a.A_enc[(size_t)(r0 + r) * a.lda + n] : 0.f;
        zv[j] = (ok && n < Z) ? a.z[(size_t)(r0 + r) * Z + n] : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        int i = t + kMidThreads * j, r = i >> 6, n = i & 63;
        float g = 0.f, v = 0.f;
        if (r0 + r < a.B) {
          float av = ao[j];
          float dy = (av * sc_o[n] + sh_o[n] > 0.f) ? go[j] * dropout_mult(a.ns_d0.drop, (uint32_t)(r0 + r), (uint32_t)n) : 0.f;
          float xh = (av - mean_o[n]) * rstd_o[n];
          g = gsc[n] * (dy - m1[n] - xh * m2[n]);
          v = hi[j] * sc_i[n] + sh_i[n];
          if (a.ns_enc.mode != NORM_RAW) v = fmaxf(v, 0.f) * dropout_mult(a.ns_enc.drop, (uint32_t)(r0 + r), (uint32_t)n);
        }
        Gn[r * kTS + n] = g; GT[n * kTS + r] = g;
        Hn[r * kTS + n] = v; An[r * kTS + n] = hi[j];
        if (n < zs) Zn[r * zs + n] = zv[j];
      }
    }
    __syncthreads();
    // dW_dec0[n][j] += sum_r G[r][n] z[r][j] ;  dz[r][j] = sum_n G[r][n] W_dec0[n][j]
    if (narrow) {
      tile_mma_n16(Gn, Zn, zs, kTileR, t, nW0);        // dW_dec0[4 ty + i][tx]
      float dzn[4] = {0.f, 0.f, 0.f, 0.f};
      tile_mma_n16(GT, Wd0n, zs, kH, t, dzn);          // dz[4 ty + i][tx]
      __syncthreads();                     // everyone is done reading z and G
#pragma unroll
      for (int i = 0; i < 4; ++i) Zn[(4 * ty + i) * zs + tx] = dzn[i];
    } else {
      tile_mma(Gn, Zn, kTileR, ty, tx, accW0);
      float dz[4][4];
      zero16(dz);
      tile_mma(GT, Wd0n, kH, ty, tx, dz);
      __syncthreads();                     // everyone is done reading z and G
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) Zn[(4 * ty + i) * kTS + 4 * tx + j] = dz[i][j];
    }
    __syncthreads();
    // latent backward, one thread per row: dz -> dPL = (d loc | d scale_raw); overwrites G with dPL
    for (int i = t; i < kTileR * kH; i += kMidThreads) { Gn[(i >> 6) * kTS + (i & 63)] = 0.f; GT[(i & 63) * kTS + (i >> 6)] = 0.f; }
    __syncthreads();
    if (t < kTileR) {
      const int r = t, b = r0 + r;
      if (b < a.B) {
        float4 nz4 = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int j = 0; j < Z; ++j) {
          const float dzz = Zn[r * zs + j];
          if (a.deterministic) {
            float g = (a.deterministic == 2 || a.PL[(size_t)b * Z + j] > 0.f) ? dzz : 0.f;
            Gn[r * kTS + j] = g; GT[j * kTS + r] = g;
          } else {
            float mu = a.loc[(size_t)b * Z + j], sg = a.scale[(size_t)b * Z + j];
            float v, dv;
            activation(a.scale_act, a.PL[(size_t)b * 2 * Z + Z + j], v, dv);
            float dmu = dzz + a.kl_weight * mu;
            float e;
            if (a.eps_z) {
              e = a.eps_z[(size_t)b * Z + j];
            } else {
              if ((j & 3) == 0) nz4 = philox_normal4(a.noise, (uint32_t)b, (uint32_t)(j >> 2), kNoiseStreamZ);
              const int k = j & 3;
              e = k == 0 ? nz4.x : (k == 1 ? nz4.y : (k == 2 ? nz4.z : nz4.w));
            }
            float dsr = (dzz * e + a.kl_weight * (sg - 1.f / sg)) * dv;
            Gn[r * kTS + j] = dmu; GT[j * kTS + r] = dmu;
            Gn[r * kTS + Z + j] = dsr; GT[(Z + j) * kTS + r] = dsr;
          }
        }
      }
    }
    __syncthreads();
    // dW_lat[m][k] += sum_r dPL[r][m] h[r][k] ; db_lat[m] += sum_r dPL[r][m] ; dH_enc[r][k] = sum_m dPL[r][m] W_lat[m][k]
    if (narrow) tile_mma_m32(Gn, Hn, kTileR, t, nW1);  // dW_lat[4 (t >> 5) + i][2 (t & 31) + j]
    else tile_mma(Gn, Hn, kTileR, ty, tx, accW1);
    if (t < ZP) {
      for (int r = 0; r < kTileR; ++r) accb += Gn[r * kTS + t];
    }
    float dh[4][4];
    zero16(dh);
    tile_mma(GT, Wln, ZP, ty, tx, dh);
    float cs[4] = {0.f, 0.f, 0.f, 0.f}, cq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int rl = 4 * ty + i, r = r0 + rl;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = 4 * tx + j;
        if (r < a.B) {
          const float v = dh[i][j];
          float dy = Hn[rl * kTS + k] > 0.f ? v * in_drop_scale : 0.f;
          float xh = (An[rl * kTS + k] - mean_i[k]) * rstd_i[k];
          cs[j] += dy; cq[j] += dy * xh;
        }
      }
      if (r < a.B) *reinterpret_cast<float4*>(a.dH_enc + (size_t)r * kH + 4 * tx) = make_float4(dh[i][0], dh[i][1], dh[i][2], dh[i][3]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) { red1[ty * kH + 4 * tx + j] = cs[j]; red2[ty * kH + 4 * tx + j] = cq[j]; }
    __syncthreads();
    if (t < kH) {
      double s1 = 0.0, s2 = 0.0;
      for (int y = 0; y < 16; ++y) { s1 += (double)red1[y * kH + t]; s2 += (double)red2[y * kH + t]; }
      dacc[t] += s1; dacc[kH + t] += s2;
    }
    __syncthreads();
  }
  if (narrow) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (tx < Z) atomicAdd(&a.dW_d0[(size_t)(4 * ty + i) * Z + tx], nW0[i]);
      const int m = 4 * (t >> 5) + i;
      if (m < ZP) {
        red_add_v2f(&a.dW_lat[(size_t)m * kH + 2 * (t & 31)], nW1[i][0], nW1[i][1]);     // rows of 64 floats, 64-float aligned tensor
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = 4 * ty + i, k = 4 * tx + j;
        if (k < Z) atomicAdd(&a.dW_d0[(size_t)n * Z + k], accW0[i][j]);
        if (n < ZP) atomicAdd(&a.dW_lat[(size_t)n * kH + k], accW1[i][j]);
      }
  }
  if (t < ZP) atomicAdd(&a.db_lat[t], accb);
  if (t < kH) {
    atomicAdd(&a.prev_sdy[t], dacc[t]);
    atomicAdd(&a.prev_sdyx[t], dacc[kH + t]);
    if (a.ns_enc.mode == NORM_BIAS) {
      atomicAdd(&a.prev_dbeta[t], (float)dacc[t]);
    } else {
      atomicAdd(&a.prev_dgamma[t], (float)dacc[kH + t]);
      atomicAdd(&a.prev_dbeta[t], (float)dacc[t]);
    }
  }
}

}  // namespace sisua
