// Per-element mathematics of the SISUA ELBO step (forward value + derivative), shared by the
// un-fused cross-check kernels and the fused tcgen05 epilogues so both paths agree by construction.
// Formulas: SURVEY.md Appendix A (restating odin-ai's NegativeBinomialDisp / ZeroInflated, which
// follow scVI's log_zinb_positive / log_nb_positive, eps = 1e-8) and sisua/models/scvi.py:117-138.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace sisua {

constexpr float kEps = 1e-8f;
constexpr float kSoftplus1Shift = 0.5413248546129181f;

enum Act : int { ACT_SOFTPLUS = 0, ACT_SOFTPLUS1 = 1, ACT_SOFTPLUS_P1 = 2, ACT_EXP = 3, ACT_IDENTITY = 4 };

// log(1 + t) for t in [0, 1]; series below 2^-6 keeps full relative accuracy where logf(1+t) cannot.
__device__ __forceinline__ float log1p_unit(float t) {
  if (t < 0.015625f) return t * (1.0f - t * (0.5f - t * (0.33333334f - 0.25f * t)));
  return __logf(1.0f + t);
}

// softplus(x) and sigmoid(x) from one exponential
__device__ __forceinline__ void softplus_sigmoid(float x, float& sp, float& sg) {
  float e = __expf(-fabsf(x));
  float r = __frcp_rn(1.0f + e);
  sp = fmaxf(x, 0.0f) + log1p_unit(e);
  sg = (x >= 0.0f) ? r : e * r;
}

__device__ __forceinline__ float softplus_only(float x) {
  float e = __expf(-fabsf(x));
  return fmaxf(x, 0.0f) + log1p_unit(e);
}

// value v = act(x) and dv = d act / dx
__device__ __forceinline__ void activation(int kind, float x, float& v, float& dv) {
  switch (kind) {
    case ACT_SOFTPLUS: softplus_sigmoid(x, v, dv); break;
    case ACT_SOFTPLUS1: softplus_sigmoid(x + kSoftplus1Shift, v, dv); break;
    case ACT_SOFTPLUS_P1: softplus_sigmoid(x, v, dv); v += 1.0f; break;
    case ACT_EXP: v = __expf(x); dv = v; break;
    default: v = x; dv = 1.0f; break;
  }
}

// psi(t), t > 0
__device__ __forceinline__ float digamma_pos(float t) {
  float r = 0.0f;
#pragma unroll 1
  while (t < 6.0f) { r -= __frcp_rn(t); t += 1.0f; }
  float it = __frcp_rn(t), it2 = it * it;
  return r + __logf(t) - 0.5f * it - it2 * (0.083333336f - it2 * (0.008333334f - it2 * 0.003968254f));
}

// lg = lgamma(x + th) - lgamma(th) - lgamma(x + 1),  dg = psi(x + th) - psi(th);  x > 0.
// Small integer counts (the bulk of non-zero entries) use the exact product form
//   Gamma(x+th) / (Gamma(th) x!) = prod_{k<x} (th + k)/(k + 1)
// -> one log and one division instead of three lgamma + two digamma.
__device__ __forceinline__ void nb_gamma_terms(float x, float th, bool want_grad, float& lg, float& dg) {
  float xr = rintf(x);
  if (x == xr && x <= 8.0f && th < 1.0e4f) {
    float q = th, dq = 1.0f, f = 1.0f;   // q = prod (th+k), dq = dq/dth, f = x!
#pragma unroll
    for (int k = 1; k < 8; ++k) {
      if ((float)k < x) {
        float tk = th + (float)k;
        dq = fmaf(dq, tk, q);
        q *= tk;
        f *= (float)(k + 1);
      }
    }
    lg = __logf(q) - __logf(f);
    dg = want_grad ? __fdividef(dq, q) : 0.0f;
  } else {
    lg = lgammaf(x + th) - lgammaf(th) - lgammaf(x + 1.0f);
    dg = want_grad ? (digamma_pos(x + th) - digamma_pos(th)) : 0.0f;
  }
}

struct CountGrad {  // d llk / d(mu, theta, pi_logit)
  float dmu, dth, dpi;
};

// Zero-inflated NB (mean / inverse dispersion / dropout logit). Returns log-likelihood of one entry.
// kNoEps: TFP's NegativeBinomial form (no 1e-8 inside the logarithms)
template <bool kZeroInflated, bool kGrad, bool kNoEps = false>
__device__ __forceinline__ float count_llk(float x, float mu, float th, float pi, CountGrad& g) {
  const float kEps = kNoEps ? 1e-30f : sisua::kEps;
  float tm = th + mu + kEps;
  float lt = __logf(th + kEps);
  float ltm = __logf(tm);
  float r_tm = __frcp_rn(tm);
  float dlog = lt - ltm;
  float n0 = th * dlog;   // log NB(0)
  float dn0_dmu = 0.f, dn0_dth = 0.f;
  if (kGrad) {
    dn0_dmu = -th * r_tm;
    dn0_dth = dlog + th * (__frcp_rn(th + kEps) - r_tm);
  }
  float llk;
  if (x < sisua::kEps) {
    if (kZeroInflated) {
      float sp_a, sg_a, sp_b, sg_b;
      softplus_sigmoid(n0 - pi, sp_a, sg_a);   // sg_a = w = sigmoid(n0 - pi)
      softplus_sigmoid(-pi, sp_b, sg_b);
      llk = sp_a - sp_b;
      if (kGrad) { g.dpi = sg_b - sg_a; g.dmu = sg_a * dn0_dmu; g.dth = sg_a * dn0_dth; }
    } else {
      llk = n0;
      if (kGrad) { g.dpi = 0.f; g.dmu = dn0_dmu; g.dth = dn0_dth; }
    }
  } else {
    float lm = __logf(mu + kEps);
    float lg, dg;
    nb_gamma_terms(x, th, kGrad, lg, dg);
    llk = n0 + x * (lm - ltm) + lg;
    if (kGrad) {
      g.dmu = dn0_dmu + x * (__frcp_rn(mu + kEps) - r_tm);
      g.dth = dn0_dth - x * r_tm + dg;
      g.dpi = 0.f;
    }
    if (kZeroInflated) {
      float sp_p, sg_p;
      softplus_sigmoid(pi, sp_p, sg_p);
      llk -= sp_p;                     // log sigmoid(-pi) = -softplus(pi)
      if (kGrad) g.dpi = -sg_p;
    }
  }
  return llk;
}

// TFP NegativeBinomial(total_count = exp(a), logits = b) for real-valued y (protein head,
// configs/base.yaml:38-40). Returns log-prob; da, db = d/d(a, b).
template <bool kGrad>
__device__ __forceinline__ float nb_tfp_llk(float y, float a, float b, float& da, float& db) {
  float r = __expf(a);
  float sp_b, sg_b;
  softplus_sigmoid(b, sp_b, sg_b);          // log sigmoid(b) = b - sp_b ; log sigmoid(-b) = -sp_b
  float llk = lgammaf(r + y) - lgammaf(r) - lgammaf(y + 1.0f) - r * sp_b + y * (b - sp_b);
  if (kGrad) {
    float dr = digamma_pos(r + y) - digamma_pos(r) - sp_b;
    da = dr * r;
    db = y - (r + y) * sg_b;
  }
  return llk;
}

// ---------------------------------------------------------------------------------------------------
// Lean evaluation for the fused epilogue with the default links (mean = softplus, dispersion =
// softplus(. + log(e-1))): straight-line MUFU ex2 / lg2 / rcp code, no per-element branches except a
// warp-uniform skip of the x > 0 terms.  Same mathematics as count_llk above (eps placement differs by
// O(1e-8/theta), far below the 1e-4 parity tolerance).
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ float mufu_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

// softplus(r) and sigmoid(r) (= its derivative)
__device__ __forceinline__ void softplus_fast(float r, float& v, float& dv) {
  float e = mufu_ex2(fminf(r, 30.f) * kLog2e);
  float s = 1.f + e;
  float big = kLn2 * mufu_lg2(s);
  float small = e * (1.f - e * (0.5f - e * (0.33333334f - 0.25f * e)));
  v = e < 0.015625f ? small : big;
  v = r > 30.f ? r : v;
  dv = e * mufu_rcp(s);
}

// Stirling forms, z >= 8
__device__ __forceinline__ float lgamma_stirling(float z, float lnz, float iz) {
  float iz2 = iz * iz;
  return (z - 0.5f) * lnz - z + 0.9189385332f + iz * (0.083333336f - iz2 * (0.0027777778f - iz2 * 0.00079365079f));
}
__device__ __forceinline__ float digamma_stirling(float lnz, float iz) {
  float iz2 = iz * iz;
  return lnz - 0.5f * iz - iz2 * (0.083333336f - iz2 * (0.008333334f - iz2 * 0.003968254f));
}

// Large (> 8) or non-integer counts: Stirling at x+th, x+1 and th+8 (all >= 8) with
// lgamma(th) = lgamma(th+8) - log prod_{k<8}(th+k)  (lnq, dgq = log and log-derivative of that product).
__device__ __noinline__ void gamma_terms_large(float x, float th, float lnq, float dgq, bool grad, float* lg, float* dg) {
  if (x >= 7.f && th < 1e4f) {
    float z1 = x + th, z2 = x + 1.f, z3 = th + 8.f;
    float l1 = kLn2 * mufu_lg2(z1), l2 = kLn2 * mufu_lg2(z2), l3 = kLn2 * mufu_lg2(z3);
    float i1 = mufu_rcp(z1), i2 = mufu_rcp(z2), i3 = mufu_rcp(z3);
    *lg = lgamma_stirling(z1, l1, i1) - lgamma_stirling(z2, l2, i2) - lgamma_stirling(z3, l3, i3) + lnq;
    *dg = digamma_stirling(l1, i1) - digamma_stirling(l3, i3) + dgq;
  } else {
    *lg = lgammaf(x + th) - lgammaf(th) - lgammaf(x + 1.f);
    *dg = grad ? digamma_pos(x + th) - digamma_pos(th) : 0.f;
  }
}

struct ElemResult { float llk, ga, gb, gl, mu, th; };   // ga, gb, gl = d llk / d raw head outputs
struct CoreResult { float llk, gmu, gth, gl; };         // d llk / d (mean, dispersion, dropout logit)

__constant__ float c_inv_int[9] = {1.f, 1.f / 2.f, 1.f / 3.f, 1.f / 4.f, 1.f / 5.f, 1.f / 6.f, 1.f / 7.f, 1.f / 8.f, 1.f / 9.f};   // [k] = 1 / (k+1)

// (ZI)NB log-likelihood of U counts given the positive parameters, and the partial derivatives.
// The U elements advance in lock-step through every warp-uniform branch so that the scheduler can interleave
// their dependent MUFU / FMA chains inside each basic block (one element at a time leaves the issue slots idle).
template <bool kZeroInflated, bool kGrad, int U>
__device__ __forceinline__ void count_core_fast(const float (&mu)[U], const float (&th)[U], const float (&pi)[U],
                                                const float (&x)[U], CoreResult (&o)[U], const float eps = kEps) {
  float Rt[U], n0[U], dn0_dmu[U], dn0_dth[U], Ep[U], Rp[U];
  bool nz[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    Rt[u] = mufu_rcp(th[u] + mu[u] + eps);
    const float rho = th[u] * Rt[u];
    const float dlog = kLn2 * mufu_lg2(rho + 1e-30f);
    n0[u] = th[u] * dlog;
    dn0_dmu[u] = -rho; dn0_dth[u] = dlog + 1.f - rho;
    Ep[u] = 0.f; Rp[u] = 1.f;
    if (kZeroInflated) {
      const float pc = fminf(fmaxf(pi[u], -60.f), 60.f);
      Ep[u] = mufu_ex2(-pc * kLog2e);            // exp(-pi)
      Rp[u] = mufu_rcp(1.f + Ep[u]);             // sigmoid(pi)
      float Eu = mufu_ex2((n0[u] - pc) * kLog2e);
      float Su = 1.f + Eu;
      float w = Eu * mufu_rcp(Su);               // sigmoid(n0 - pi)
      o[u].llk = kLn2 * mufu_lg2(Su * Rp[u]);    // softplus(n0 - pi) - softplus(-pi)
      o[u].gl = Ep[u] * Rp[u] - w; o[u].gmu = w * dn0_dmu[u]; o[u].gth = w * dn0_dth[u];
    } else {
      o[u].llk = n0[u]; o[u].gmu = dn0_dmu[u]; o[u].gth = dn0_dth[u]; o[u].gl = 0.f;
    }
    nz[u] = x[u] >= kEps;
  }
  // per-lane trip count of the rising-factorial product, as a float (no int<->float conversions: those run on the
  // same quarter-rate unit as the transcendentals); non-negative floats order like their bit patterns
  float lim[U];
  bool small_x[U];
  float lim_max = 0.f;
  bool any_big = false;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const float xr = (x[u] + 8388608.f) - 8388608.f;          // rint(x) for 0 <= x < 2^22
    // the product form needs (th+7)^8 inside fp32 range; beyond that the rare out-of-line path takes over
    small_x[u] = (x[u] == xr) && x[u] <= 8.f && th[u] < 1e4f;
    lim[u] = nz[u] ? (small_x[u] ? x[u] : 8.f) : 0.f;
    lim_max = fmaxf(lim_max, lim[u]);
    any_big = any_big || (nz[u] && !small_x[u]);
  }
  const float kmax = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(lim_max)));
  if (kmax > 0.f) {            // some lane of the warp holds a non-zero count
    // lg = lgamma(x+th) - lgamma(th) - lgamma(x+1), dg = psi(x+th) - psi(th)
    float q[U], dq[U], finv[U];          // q = prod_{k<min(x,8)} (th+k), finv = 1 / min(x,8)!
#pragma unroll
    for (int u = 0; u < U; ++u) { q[u] = th[u]; dq[u] = 1.f; finv[u] = 1.f; }
    // the 32 cells of a warp rarely hold more than a few counts at one gene: the loop stops at the warp's largest
    float kf = 1.f;
#pragma unroll 1
    for (int k = 1; kf < kmax; ++k, kf += 1.f) {
      const float inv = c_inv_int[k];     // 1 / (k+1), uniform constant load
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (kf < lim[u]) {
          float tk = th[u] + kf;
          dq[u] = fmaf(dq[u], tk, q[u]);
          q[u] *= tk;
          finv[u] *= inv;
        }
      }
    }
    float lg[U], dg[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      lg[u] = kLn2 * mufu_lg2(q[u] * finv[u]);
      dg[u] = dq[u] * mufu_rcp(q[u]);
    }
    if (__any_sync(0xffffffffu, any_big)) {
#pragma unroll
      for (int u = 0; u < U; ++u) {                                     // out of line: rare, keeps the hot loop small
        float lg_big, dg_big;
        gamma_terms_large(x[u], th[u], kLn2 * mufu_lg2(q[u]), dg[u], kGrad, &lg_big, &dg_big);
        const bool take = nz[u] && !small_x[u];
        lg[u] = take ? lg_big : lg[u];
        dg[u] = take ? dg_big : dg[u];
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float lnm = kLn2 * mufu_lg2((mu[u] + eps) * Rt[u]);
      float llk1 = n0[u] + x[u] * lnm + lg[u];
      float gmu1 = dn0_dmu[u] + x[u] * (mufu_rcp(mu[u] + eps) - Rt[u]);
      float gth1 = dn0_dth[u] - x[u] * Rt[u] + dg[u];
      float gl1 = 0.f;
      if (kZeroInflated) {
        llk1 += kLn2 * mufu_lg2(Ep[u] * Rp[u]) - fmaxf(pi[u] - 60.f, 0.f);   // log sigmoid(-pi)
        gl1 = -Rp[u];
      }
      o[u].llk = nz[u] ? llk1 : o[u].llk; o[u].gmu = nz[u] ? gmu1 : o[u].gmu;
      o[u].gth = nz[u] ? gth1 : o[u].gth; o[u].gl = nz[u] ? gl1 : o[u].gl;
    }
  }
}

// scalar fall-back of the pair-wise epilogue mathematics (pair_math.cuh) for negative / NaN inputs (not count data); every
// lane of the warp calls it together (count_core_fast votes warp-wide)
namespace pm {
template <bool ZI, bool GRAD>
__device__ __noinline__ void core_scalar_fallback(float mu, float th, float pi, float x, float eps, float& llk, float& gmu, float& gth,
                                                  float& gl) {
  const float mu1[1] = {mu}, th1[1] = {th}, pi1[1] = {pi}, x1[1] = {x};
  CoreResult c[1];
  count_core_fast<ZI, GRAD, 1>(mu1, th1, pi1, x1, c, eps);
  llk = c[0].llk; gmu = c[0].gmu; gth = c[0].gth; gl = c[0].gl;
}
}  // namespace pm

// default links of the VAE / DCA / SISUA heads: mean = softplus(ra), dispersion = softplus(rb + log(e-1))
template <bool kZeroInflated, bool kGrad, int U>
__device__ __forceinline__ void count_elem_fast(const float (&ra)[U], const float (&rb)[U], const float (&pi)[U],
                                                const float (&x)[U], ElemResult (&o)[U]) {
  float mu[U], dmu[U], th[U], dth[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    softplus_fast(ra[u], mu[u], dmu[u]);
    softplus_fast(rb[u] + kSoftplus1Shift, th[u], dth[u]);
  }
  CoreResult c[U];
  count_core_fast<kZeroInflated, kGrad, U>(mu, th, pi, x, c);
#pragma unroll
  for (int u = 0; u < U; ++u) {
    o[u].mu = mu[u]; o[u].th = th[u]; o[u].llk = c[u].llk;
    o[u].ga = c[u].gmu * dmu[u]; o[u].gb = c[u].gth * dth[u]; o[u].gl = c[u].gl;
  }
}

// scVI links (scvi.py:64-86): mean = exp(clip(library)) * clamp(softmax_g(u)), dispersion = exp(rb).
//   u_lse = u_g - logsumexp_g(u), eL = exp(clipped library).
//   t = d llk / d s_raw_g (softmax output before the clamp); the softmax Jacobian s_raw (t - sum_j s_raw_j t_j)
//   is finished by the caller once the row sum is known.  gmu_mu = (d llk / d mean) * mean -> d llk / d library.
struct ScviElem { float llk, mu, th, s_raw, t, gmu_mu, gb, gl; };

template <bool kZeroInflated, bool kGrad, int U>
__device__ __forceinline__ void count_elem_scvi(const float (&u_lse)[U], const float (&rb)[U], const float (&pi)[U],
                                                const float (&x)[U], float eL, ScviElem (&o)[U]) {
  const float lo = 1e-7f, hi = 1.f - 1e-7f;
  float mu[U], th[U];
  bool inside[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const float s_raw = mufu_ex2(u_lse[u] * kLog2e);
    inside[u] = s_raw >= lo && s_raw <= hi;
    o[u].s_raw = s_raw;
    mu[u] = eL * fminf(fmaxf(s_raw, lo), hi);
    th[u] = mufu_ex2(rb[u] * kLog2e);
  }
  CoreResult c[U];
  count_core_fast<kZeroInflated, kGrad, U>(mu, th, pi, x, c);
#pragma unroll
  for (int u = 0; u < U; ++u) {
    o[u].mu = mu[u]; o[u].th = th[u]; o[u].llk = c[u].llk;
    o[u].t = inside[u] ? c[u].gmu * eL : 0.f;
    o[u].gmu_mu = c[u].gmu * mu[u];
    o[u].gb = c[u].gth * th[u];
    o[u].gl = c[u].gl;
  }
}

// Counter-based dropout masks (Philox4x32-10, counter = (row, col/8, step, stream)): a pure function, so
// forward and backward regenerate it instead of storing it, and the CPU oracle reproduces it exactly
// (oracle/philox.py).  stream 0 = input dropout on log1p(x); stream 1+u = hidden unit u.
struct DropSpec {
  float rate;      // 0 -> disabled
  float scale;     // 1 / (1 - rate)
  uint32_t seed_lo, seed_hi, step, stream;
  const long long* step_ptr;   // non-null: step = *step_ptr + 1 (device-side optimiser step counter; CUDA-graph replays)
};

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u; k.y += 0xBB67AE85u;
  }
  return c;
}

// EIGHT consecutive columns (col8*8 .. col8*8+7) of one row from ONE Philox call: column j of the group uses the
// (j&1 ? high : low) 16 bits of output word j>>1 as a uniform in [0, 65536); kept when it is >= floor(rate * 65536).
// multiplier is 0 or 1/(1-rate).  Out of line on purpose: it sits behind a `rate > 0` test at every use, and inlining
// ten Philox rounds into the unrolled tile loaders of the small kernels multiplied their code size.  The spec is
// passed by value: a reference would force the caller's by-value argument struct into local memory.
struct DropMult8 { float m[8]; };
__device__ __forceinline__ uint32_t dropout_step(const DropSpec& d) {
  return d.step_ptr ? (uint32_t)(*d.step_ptr + 1) : d.step;
}
// inline form for the streaming first-layer kernel (step resolved once by the caller)
__device__ __forceinline__ void dropout_mult8_inline(const DropSpec& d, uint32_t step, uint32_t row, uint32_t col8, float* m) {
  uint4 r = philox4x32_10(make_uint4(row, col8, step, d.stream), make_uint2(d.seed_lo, d.seed_hi));
  const uint32_t thr = (uint32_t)(d.rate * 65536.0f), thr_hi = thr << 16;
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    m[2 * j] = (w[j] & 0xffffu) >= thr ? d.scale : 0.f;
    m[2 * j + 1] = w[j] >= thr_hi ? d.scale : 0.f;        // (w >> 16) >= thr
  }
}
__device__ __noinline__ DropMult8 dropout_mult8(DropSpec d, uint32_t row, uint32_t col8) {
  DropMult8 o;
  dropout_mult8_inline(d, dropout_step(d), row, col8, o.m);
  return o;
}
__device__ __forceinline__ float dropout_mult(const DropSpec& d, uint32_t row, uint32_t col) {
  if (d.rate <= 0.f) return 1.f;
  DropMult8 m = dropout_mult8(d, row, col >> 3);
  return m.m[col & 7u];
}

// Reparameterisation noise drawn in-kernel (eps == NULL -> Philox): standard normals as a pure
// function of (seed; row, column group of 4, step, stream), Box-Muller on the four words of one Philox4x32-10 call:
//   u1 = ((w0 >> 8) + 0.5) 2^-24 in (0, 1), u2 = (w1 >> 8) 2^-24 in [0, 1): n0 = sqrt(-2 ln u1) cos(2 pi u2), n1 = .. sin ..
// and the same on (w2, w3).  Forward and backward regenerate it; oracle/philox.py:normal_noise reproduces it.
// stream = kNoiseStreamZ + 2 s for the latent of Monte-Carlo sample s, kNoiseStreamL + 2 s for scVI's library latent.
constexpr uint32_t kNoiseStreamZ = 0x100u, kNoiseStreamL = 0x101u;
struct NoiseSpec {
  uint32_t seed_lo, seed_hi, step;
  const long long* step_ptr;   // non-null: step = *step_ptr + 1 (device-side optimiser step counter; CUDA-graph replays)
};
__device__ __noinline__ float4 philox_normal4(NoiseSpec n, uint32_t row, uint32_t group, uint32_t stream) {
  const uint32_t step = n.step_ptr ? (uint32_t)(*n.step_ptr + 1) : n.step;
  const uint4 r = philox4x32_10(make_uint4(row, group, step, stream), make_uint2(n.seed_lo, n.seed_hi));
  const float k24 = 5.9604644775390625e-08f;     // 2^-24
  const float ra = sqrtf(-2.0f * logf(((float)(r.x >> 8) + 0.5f) * k24));
  const float rb = sqrtf(-2.0f * logf(((float)(r.z >> 8) + 0.5f) * k24));
  float sa, ca, sb, cb;
  sincospif(2.0f * ((float)(r.y >> 8) * k24), &sa, &ca);
  sincospif(2.0f * ((float)(r.w >> 8) * k24), &sb, &cb);
  return make_float4(ra * ca, ra * sa, rb * cb, rb * sb);
}
// element `col` of the noise row (one Philox call per element: for the scalar library latent and ragged accesses)
__device__ __forceinline__ float philox_normal(NoiseSpec n, uint32_t row, uint32_t col, uint32_t stream) {
  const float4 q = philox_normal4(n, row, col >> 2, stream);
  const uint32_t k = col & 3u;
  return k == 0 ? q.x : (k == 1 ? q.y : (k == 2 ? q.z : q.w));
}

// Programmatic dependent launch (sm_90+): let the next kernel on the stream be scheduled early / wait until every
// predecessor grid has completed and flushed.  Both are no-ops for a normally launched kernel.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// vector reductions into global memory (sm_90+): one L2 atomic operation for 2 / 4 consecutive floats
__device__ __forceinline__ void red_add_v4f(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_add_v2f(float* addr, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}

// warp / block reductions -------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// block-wide sum; every thread gets the result. `scratch` needs >= 33 elements.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* scratch) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[w] = v;
  __syncthreads();
  if (w == 0) {
    T t = (lane < nw) ? scratch[lane] : T(0);
    t = warp_sum(t);
    if (lane == 0) scratch[32] = t;
  }
  __syncthreads();
  return scratch[32];
}
__device__ __forceinline__ float block_max(float v, float* scratch) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) scratch[w] = v;
  __syncthreads();
  if (w == 0) {
    float t = (lane < nw) ? scratch[lane] : -INFINITY;
    t = warp_max(t);
    if (lane == 0) scratch[32] = t;
  }
  __syncthreads();
  return scratch[32];
}

}  // namespace sisua
