// Pair-wise (two genes per thread, lock-step) evaluation of the (ZI)NB count log-likelihood and its derivatives for the
// fused output-head epilogue.  Same mathematics as device_math.cuh:count_llk (odin-ai's
// NegativeBinomialDisp / ZeroInflated following scVI's log_zinb_positive, eps = 1e-8), arranged for Blackwell's issue
// and special-function limits, which are what bound that epilogue (profiles/README.md):
//   * both genes advance through one straight-line instruction stream, so every add / mul / fma is ONE packed
//     FADD2 / FMUL2 / FFMA2 (sm_100 fp32x2 instructions) instead of two scalar ones: half the issue slots;
//   * reciprocals are merged (one MUFU.RCP of a product, two multiplies to take it apart), logarithms of products are
//     taken once, and the small-count rising factorial is branch-free, so a zero count costs 11 MUFU operations and a
//     count in 1..3 costs 3 more (18 before);
//   * selected exponentials run on the FMA pipe as a degree-5 polynomial (Cody-Waite split, packed Horner steps) to
//     balance the MUFU and FMA pipes;
//   * counts above 3 (a quarter of the warp x gene-pair groups of a pbmc-like matrix hold one) stay packed and branch-free
//     through Stirling's series with a shift-by-4 recurrence (core_big); only negative / NaN inputs (not count
//     data) fall back to the scalar routines of device_math.cuh.
// Compiles for the host as well (plain float arithmetic in place of the packed / MUFU instructions): tests/csrc/ builds
// it with g++ and checks the formulas against a float64 restatement without a GPU.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define PM_HD __device__ __forceinline__
#define PM_DEVICE_CODE 1
#else
#define PM_HD inline
#define PM_DEVICE_CODE 0
#endif

namespace sisua {
namespace pm {

struct F2 { float x, y; };

constexpr float kEps = 1e-8f;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kSoftplus1Shift = 0.5413248546129181f;
constexpr float kLinkClamp = 40.f;      // exp(40) = 2^57.7: products of two such terms stay inside fp32
constexpr float kTiny = 1.2e-38f;

PM_HD F2 mk(float a, float b) { F2 r; r.x = a; r.y = b; return r; }
PM_HD F2 bc(float a) { return mk(a, a); }

#if PM_DEVICE_CODE
__device__ __forceinline__ float2 as2(F2 a) { return make_float2(a.x, a.y); }
__device__ __forceinline__ F2 fr2(float2 a) { return mk(a.x, a.y); }
#ifdef SISUA_PM_SCALAR      // experiment: two scalar instructions instead of one packed fp32x2 instruction
__device__ __forceinline__ F2 add(F2 a, F2 b) { return mk(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ F2 mul(F2 a, F2 b) { return mk(a.x * b.x, a.y * b.y); }
__device__ __forceinline__ F2 fma2(F2 a, F2 b, F2 c) { return mk(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
#else
__device__ __forceinline__ F2 add(F2 a, F2 b) { return fr2(__fadd2_rn(as2(a), as2(b))); }
__device__ __forceinline__ F2 mul(F2 a, F2 b) { return fr2(__fmul2_rn(as2(a), as2(b))); }
__device__ __forceinline__ F2 fma2(F2 a, F2 b, F2 c) { return fr2(__ffma2_rn(as2(a), as2(b), as2(c))); }
#endif
__device__ __forceinline__ float ex2s(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2s(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcps(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sat(float x) { return __saturatef(x); }
__device__ __forceinline__ bool any_lane(bool p) { return __any_sync(0xffffffffu, p); }
#else
inline F2 add(F2 a, F2 b) { return mk(a.x + b.x, a.y + b.y); }
inline F2 mul(F2 a, F2 b) { return mk(a.x * b.x, a.y * b.y); }
inline F2 fma2(F2 a, F2 b, F2 c) { return mk(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
inline float ex2s(float x) { return exp2f(x); }
inline float lg2s(float x) { return log2f(x); }
inline float rcps(float x) { return 1.0f / x; }
inline float sat(float x) { return x < 0.f ? 0.f : (x > 1.f ? 1.f : x); }
inline bool any_lane(bool p) { return p; }
#endif

PM_HD F2 sub(F2 a, F2 b) { return add(a, mk(-b.x, -b.y)); }
PM_HD F2 neg(F2 a) { return mk(-a.x, -a.y); }
PM_HD F2 ex2(F2 a) { return mk(ex2s(a.x), ex2s(a.y)); }
PM_HD F2 lg2(F2 a) { return mk(lg2s(a.x), lg2s(a.y)); }
PM_HD F2 rcp(F2 a) { return mk(rcps(a.x), rcps(a.y)); }
PM_HD F2 min2(F2 a, float m) { return mk(fminf(a.x, m), fminf(a.y, m)); }
PM_HD F2 max2(F2 a, float m) { return mk(fmaxf(a.x, m), fmaxf(a.y, m)); }

PM_HD uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
PM_HD float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

// 2^t on the FMA pipe: t = n + f, n = rint(t), f in [-1/2, 1/2]; 2^f by a degree-5 polynomial (max relative error 2.4e-7
// in fp32), 2^n by adding n to the exponent field.  t is clamped below at -125 (result 2^-125 instead of a denormal).
PM_HD F2 ex2_poly(F2 t) {
  const float kMagic = 12582912.f;     // 1.5 * 2^23: adding it leaves rint(t) in the low mantissa bits
  t = max2(t, -125.f);
  const F2 j = add(t, bc(kMagic));
  const F2 f = sub(t, sub(j, bc(kMagic)));
  F2 p = fma2(f, bc(0.0013276468962430954f), bc(0.009675540961325169f));
  p = fma2(p, f, bc(0.05550713464617729f));
  p = fma2(p, f, bc(0.24022120237350464f));
  p = fma2(p, f, bc(0.6931469440460205f));
  p = fma2(p, f, bc(1.0000001192092896f));
  // (bits(j) << 23) keeps exactly n << 23: the magic constant's own bits shift out of the word
  return mk(u2f(f2u(p.x) + (f2u(j.x) << 23)), u2f(f2u(p.y) + (f2u(j.y) << 23)));
}

template <bool POLY>
PM_HD F2 ex2_sel(F2 t) { return POLY ? ex2_poly(t) : ex2(t); }

// ---- which exponentials leave the MUFU pipe (tuned on the B200: see profiles/) ----
#ifndef SISUA_PM_POLY_LINKS
#define SISUA_PM_POLY_LINKS 0     // the two softplus links (measured: 0.404 ms with, 0.396 ms without -- the FMA pipe is the busier one)
#endif
#ifndef SISUA_PM_POLY_PI
#define SISUA_PM_POLY_PI 0        // exp(-pi) of the dropout logit
#endif

// softplus(r) and its derivative sigmoid(r) for both lanes: v = log(1 + e^r), dv = e^r / (1 + e^r).
//   m = min(r, 40); e = 2^(m log2 e); s = 1 + e; v = ln2 lg2(s) + (e - (s - 1)) / s + (r - m)
// The middle term restores the bits the rounding of 1 + e loses (full relative accuracy for very negative r without a
// branch); (r - m) is exact 0 below the clamp and makes v = r above it.  `rs` = 1 / s is supplied by the caller, who
// merges the reciprocals of both links into one MUFU operation.
struct Link { F2 e, s, L; };
PM_HD Link link_begin(F2 r) {
  Link k;
  const F2 m = min2(r, kLinkClamp);
  k.e = ex2_sel<SISUA_PM_POLY_LINKS != 0>(mul(m, bc(kLog2e)));
  k.s = add(k.e, bc(1.f));
  k.L = lg2(k.s);
  return k;
}
PM_HD F2 link_value(const Link& k, F2 r, F2 rs) {
  const F2 resid = sub(k.e, sub(k.s, bc(1.f)));
  const F2 over = sub(r, min2(r, kLinkClamp));
  return fma2(k.L, bc(kLn2), fma2(resid, rs, over));
}
// value only (inference): the correction uses (2 - s) in place of 1 / s, exact enough where it matters (e << 1)
PM_HD F2 link_value_nograd(const Link& k, F2 r) {
  const F2 resid = sub(k.e, sub(k.s, bc(1.f)));
  const F2 over = sub(r, min2(r, kLinkClamp));
  const F2 w = max2(sub(bc(2.f), k.s), 0.f);
  return fma2(k.L, bc(kLn2), fma2(resid, w, over));
}

struct Core { F2 llk, gmu, gth, gl; };     // natural-log units; d llk / d (mean, inverse dispersion, dropout logit)

// Scalar fall-back for one element: sisua::pm::core_scalar_fallback<ZI, GRAD>(mu, th, pi, x, eps, llk&, gmu&, gth&, gl&) must be
// declared BEFORE this header is included -- device_math.cuh does (out of line, on top of count_core_fast<.., 1>); the
// host harness restates it in double precision.

// (ZI)NB log-likelihood of two counts given positive (mean, inverse dispersion) and the dropout logit, in three pieces so
// that a caller can run SEVERAL pairs through each straight-line piece back to back: the compiler then interleaves their
// instruction streams (independent dependency chains), which is what hides the MUFU / FMA latencies -- a single pair is
// one long chain.
struct CoreState {
  F2 mu, th, pi, x, Rt, rho, n0, dn0_dth, Ep, Rp;
  float eps = kEps;      // the 1e-8 inside the logarithms of the mean / dispersion form (scVI's log_nb_positive); ~0 for TFP's form
};

// piece 1: everything a zero count needs (and the shared terms of the non-zero case)
// `nozi` (uniform): evaluate the count distribution WITHOUT its zero inflation although the head has a dropout logit --
// the "imputed" distribution of sisua/analysis/posterior.py:210-220.
template <bool ZI, bool GRAD>
PM_HD Core core_zero(CoreState& c, bool nozi = false) {
  Core o;
  const F2 tm = add(add(c.th, c.mu), bc(c.eps));
  c.Rt = rcp(tm);
  c.rho = mul(c.th, c.Rt);
  // log2(theta / (theta + mu)).  rho carries the rounding of the approximate reciprocal (~1e-7 relative), which the
  // logarithm turns into an ABSOLUTE error that n0 = theta * log(rho) then scales by theta; the exact residual
  // theta - rho * tm (one fma) restores it to first order where it matters (rho ~ 1, i.e. theta >> mu).
  const F2 res = fma2(neg(c.rho), tm, c.th);
  const F2 lr = fma2(mul(res, c.Rt), bc(kLog2e), lg2(add(c.rho, bc(1e-30f))));
  c.n0 = mul(c.th, lr);                               // log2 NB(0)
  c.dn0_dth = fma2(lr, bc(kLn2), sub(bc(1.f), c.rho));
  c.Ep = bc(0.f); c.Rp = bc(1.f);
  if (ZI && !nozi) {
    const F2 pc = max2(min2(c.pi, kLinkClamp), -kLinkClamp);
    const F2 p2 = mul(pc, bc(kLog2e));
    c.Ep = ex2_sel<SISUA_PM_POLY_PI != 0>(neg(p2));   // exp(-pi)
    const F2 Eu = ex2(sub(c.n0, p2));                 // exp(n0 - pi)
    const F2 Sp = add(c.Ep, bc(1.f)), Su = add(Eu, bc(1.f));
    if (GRAD) {
      const F2 R2 = rcp(mul(Sp, Su));
      c.Rp = mul(R2, Su);                             // sigmoid(pi)
      const F2 w = mul(Eu, mul(R2, Sp));              // sigmoid(n0 - pi)
      o.llk = mul(lg2(mul(Su, c.Rp)), bc(kLn2));      // softplus(n0 - pi) - softplus(-pi)
      o.gl = fma2(c.Ep, c.Rp, neg(w));
      o.gmu = neg(mul(w, c.rho));
      o.gth = mul(w, c.dn0_dth);
    } else {
      c.Rp = rcp(Sp);
      o.llk = mul(lg2(mul(Su, c.Rp)), bc(kLn2));
      o.gl = o.gmu = o.gth = bc(0.f);
    }
  } else {
    o.llk = mul(c.n0, bc(kLn2));
    o.gmu = neg(c.rho); o.gth = c.dn0_dth; o.gl = bc(0.f);
  }
  return o;
}

PM_HD bool pair_nonzero(F2 x) { return x.x >= kEps || x.y >= kEps; }
// counts 1..3 (the bulk of the non-zero entries) take the closed forms of piece 2a; anything else >= 1 takes the Stirling
// forms of piece 2b (valid for any positive count, integer or not); negative / NaN inputs go to the scalar routines
PM_HD bool count_big(float x) {
  const float r = (x + 8388608.f) - 8388608.f;      // rint(x) for 0 <= x < 2^22
  return x >= kEps && !(x == r && x <= 3.f);
}
PM_HD bool count_odd(float x) { return x < 0.f || !(x == x); }

// piece 2a: counts in {0, 1, 2, 3}, branch-free
template <bool ZI, bool GRAD>
PM_HD void core_small(const CoreState& c, Core& o, bool nozi = false) {
  const F2 x = c.x, th = c.th;
  // lgamma(x + th) - lgamma(th) - lgamma(x + 1) = log(q / x!),  q = th (th+1)^[x>=2] (th+2)^[x>=3]
  const F2 i2 = mk(sat(x.x - 1.f), sat(x.y - 1.f)), i3 = mk(sat(x.x - 2.f), sat(x.y - 2.f));
  const F2 t1 = add(th, bc(1.f));
  const F2 g1 = fma2(i2, th, bc(1.f)), g2 = fma2(i3, t1, bc(1.f));
  const F2 g12 = mul(g1, g2);
  const F2 q = mul(th, g12);
  const F2 mue = add(c.mu, bc(c.eps));
  const F2 m = mul(mue, c.Rt);                        // mu / (theta + mu)
  // m^x = m * (x >= 2 ? m : 1) * (x >= 3 ? m : 1); the factors are blended as i*m + (1 - i): no cancellation for tiny m
  const F2 mx = mul(m, mul(fma2(i2, m, sub(bc(1.f), i2)), fma2(i3, m, sub(bc(1.f), i3))));
  const F2 finv = fma2(fma2(x, bc(1.f / 12.f), bc(-0.75f)), x, bc(5.f / 3.f));    // 1 / x! for x = 1, 2, 3
  const F2 prod = max2(mul(mul(mx, q), finv), kTiny);
  F2 l2 = add(c.n0, lg2(prod));                       // log2 units
  if (ZI && !nozi) {
    // log sigmoid(-pi) = log2(Ep Rp) (clamped logit) - the part of pi above the clamp
    const F2 over = max2(sub(c.pi, bc(kLinkClamp)), 0.f);
    l2 = add(l2, fma2(over, bc(-kLog2e), lg2(mul(c.Ep, c.Rp))));
  }
  const F2 llk1 = mul(l2, bc(kLn2));
  const bool nz0 = x.x >= kEps, nz1 = x.y >= kEps;
  o.llk = mk(nz0 ? llk1.x : o.llk.x, nz1 ? llk1.y : o.llk.y);
  if (GRAD) {
    const F2 dq = fma2(th, fma2(i2, g2, mul(g1, i3)), g12);     // dq / dth
    const F2 R3 = rcp(mul(q, mue));
    const F2 Rq = mul(R3, mue), Rm = mul(R3, q);
    const F2 gmu1 = fma2(x, sub(Rm, c.Rt), neg(c.rho));
    const F2 gth1 = fma2(dq, Rq, fma2(neg(x), c.Rt, c.dn0_dth));
    o.gmu = mk(nz0 ? gmu1.x : o.gmu.x, nz1 ? gmu1.y : o.gmu.y);
    o.gth = mk(nz0 ? gth1.x : o.gth.x, nz1 ? gth1.y : o.gth.y);
    if (ZI) o.gl = mk(nz0 ? -c.Rp.x : o.gl.x, nz1 ? -c.Rp.y : o.gl.y);
  }
}

// piece 2b: any positive count (integer or not), still packed and branch-free.  The three log-gammas of
//   L = lgamma(x + th) - lgamma(th) - lgamma(x + 1),   D = dL/dth = psi(x + th) - psi(th)
// come from Stirling's series; an argument below 4 is shifted up by 4 through z (z+1) (z+2) (z+3) = u (u + 2),
// u = z (z + 3) (value) and its derivative (2z + 3)(2u + 2) (digamma), selected per element.  The difference
// lgamma(z2) - lgamma(z1) is arranged as (z1 - 1/2) ln(z2/z1) + (z2 - z1)(ln z2 - 1) with ln(z2/z1) taken of the
// quotient, so that th >> x (z2/z1 -> 1) loses no more than the rounding of the quotient; the series are cut at
// z^-5 (value) / z^-6 (digamma): truncation < 4e-8 / 6e-8 at z = 4.  9 MUFU operations per element on top of piece 1.
struct Shifted { F2 zs, Pe, dPe; };
template <bool GRAD>
PM_HD Shifted shift4(F2 z) {
  Shifted r;
  const bool s0 = z.x < 4.f, s1 = z.y < 4.f;
  const F2 u = mul(z, add(z, bc(3.f)));
  const F2 P = mul(u, add(u, bc(2.f)));
  r.zs = add(z, mk(s0 ? 4.f : 0.f, s1 ? 4.f : 0.f));
  r.Pe = mk(s0 ? P.x : 1.f, s1 ? P.y : 1.f);            // (selected, not blended: P can be far below 1 ulp of 1)
  r.dPe = bc(0.f);
  if (GRAD) {
    const F2 dP = mul(fma2(z, bc(2.f), bc(3.f)), fma2(u, bc(2.f), bc(2.f)));
    r.dPe = mk(s0 ? dP.x : 0.f, s1 ? dP.y : 0.f);
  }
  return r;
}
PM_HD F2 stirling_corr(F2 r) {       // 1/(12 z) - 1/(360 z^3) + 1/(1260 z^5), r = 1/z
  const F2 r2 = mul(r, r);
  return mul(r, fma2(r2, fma2(r2, bc(1.f / 1260.f), bc(-1.f / 360.f)), bc(1.f / 12.f)));
}
PM_HD F2 digamma_corr(F2 r) {        // 1/(12 z^2) - 1/(120 z^4) + 1/(252 z^6)
  const F2 r2 = mul(r, r);
  return mul(r2, fma2(r2, fma2(r2, bc(1.f / 252.f), bc(-1.f / 120.f)), bc(1.f / 12.f)));
}

template <bool ZI, bool GRAD>
PM_HD void core_big(const CoreState& c, Core& o, bool nozi = false) {
  const F2 x = c.x, th = c.th;
  const Shifted s1 = shift4<GRAD>(th), s2 = shift4<GRAD>(add(th, x)), s3 = shift4<false>(add(x, bc(1.f)));
  const F2 R12 = rcp(mul(s1.zs, s2.zs));
  const F2 r1 = mul(R12, s2.zs), r2 = mul(R12, s1.zs), r3 = rcp(s3.zs);
  const F2 lr = lg2(mul(s2.zs, r1));                      // log2(z2 / z1)
  const F2 l2s = lg2(s2.zs), l3s = lg2(s3.zs);
  const F2 dz = sub(s2.zs, s1.zs);
  // log2 of the Stirling main terms; the 1/2 log(2 pi) of z2 and z1 cancel, z3's stays
  F2 L2 = fma2(sub(s1.zs, bc(0.5f)), lr, mul(dz, sub(l2s, bc(kLog2e))));
  L2 = sub(L2, fma2(sub(s3.zs, bc(0.5f)), l3s, fma2(s3.zs, bc(-kLog2e), bc(1.3257480647361593f))));   // 1/2 log2(2 pi)
  const F2 C = sub(sub(stirling_corr(r2), stirling_corr(r1)), stirling_corr(r3));
  const F2 RP = rcp(mul(s1.Pe, s2.Pe));
  const F2 iP1 = mul(RP, s2.Pe), iP2 = mul(RP, s1.Pe);
  L2 = add(fma2(C, bc(kLog2e), L2), lg2(mul(mul(s1.Pe, s3.Pe), iP2)));
  const F2 mue = add(c.mu, bc(c.eps));
  const F2 m = max2(mul(mue, c.Rt), kTiny);             // mu / (theta + mu)
  F2 l2 = add(fma2(x, lg2(m), c.n0), L2);               // log2 units
  if (ZI && !nozi) {
    const F2 over = max2(sub(c.pi, bc(kLinkClamp)), 0.f);
    l2 = add(l2, fma2(over, bc(-kLog2e), lg2(mul(c.Ep, c.Rp))));
  }
  const F2 llk1 = mul(l2, bc(kLn2));
  const bool nz0 = x.x >= kEps, nz1 = x.y >= kEps;
  o.llk = mk(nz0 ? llk1.x : o.llk.x, nz1 ? llk1.y : o.llk.y);
  if (GRAD) {
    // psi(z) = ln zs - 1/(2 zs) - digamma_corr(1/zs) - P'/P;   r2 - r1 = -(z2s - z1s) / (z1s z2s) exactly
    F2 D = fma2(lr, bc(kLn2), mul(mul(dz, R12), bc(0.5f)));
    D = sub(D, sub(digamma_corr(r2), digamma_corr(r1)));
    D = add(D, fma2(s1.dPe, iP1, neg(mul(s2.dPe, iP2))));
    const F2 Rmue = rcp(mue);
    const F2 gmu1 = fma2(x, sub(Rmue, c.Rt), neg(c.rho));
    const F2 gth1 = add(D, fma2(neg(x), c.Rt, c.dn0_dth));
    o.gmu = mk(nz0 ? gmu1.x : o.gmu.x, nz1 ? gmu1.y : o.gmu.y);
    o.gth = mk(nz0 ? gth1.x : o.gth.x, nz1 ? gth1.y : o.gth.y);
    if (ZI) o.gl = mk(nz0 ? -c.Rp.x : o.gl.x, nz1 ? -c.Rp.y : o.gl.y);
  }
}

// piece 2c: a gene at which some cell of the warp holds a NEGATIVE (or NaN) count -- never on the reference's inputs -- is
// redone by the scalar routines of device_math.cuh (every lane of the warp calls them together)
template <bool ZI, bool GRAD>
PM_HD void core_general(const CoreState& c, Core& o, bool nozi = false) {
  float l, gm, gt, gg;
  if (any_lane(count_odd(c.x.x))) {
    if (ZI && !nozi) core_scalar_fallback<ZI, GRAD>(c.mu.x, c.th.x, c.pi.x, c.x.x, c.eps, l, gm, gt, gg);
    else core_scalar_fallback<false, GRAD>(c.mu.x, c.th.x, c.pi.x, c.x.x, c.eps, l, gm, gt, gg);
    o.llk.x = l; o.gmu.x = gm; o.gth.x = gt; o.gl.x = gg;
  }
  if (any_lane(count_odd(c.x.y))) {
    if (ZI && !nozi) core_scalar_fallback<ZI, GRAD>(c.mu.y, c.th.y, c.pi.y, c.x.y, c.eps, l, gm, gt, gg);
    else core_scalar_fallback<false, GRAD>(c.mu.y, c.th.y, c.pi.y, c.x.y, c.eps, l, gm, gt, gg);
    o.llk.y = l; o.gmu.y = gm; o.gth.y = gt; o.gl.y = gg;
  }
}

// NP pairs in lock-step: zero-count terms for all, one warp vote, then the small-count or the general path for all
template <bool ZI, bool GRAD, int NP>
PM_HD void core_multi(CoreState (&c)[NP], Core (&o)[NP], bool nozi = false) {
  bool nz = false, big = false, odd = false;
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    o[p] = core_zero<ZI, GRAD>(c[p], nozi);
    nz = nz || pair_nonzero(c[p].x); big = big || count_big(c[p].x.x) || count_big(c[p].x.y);
    odd = odd || count_odd(c[p].x.x) || count_odd(c[p].x.y);
  }
  if (!any_lane(nz)) return;                         // every cell of the warp has zeros at all these genes
  if (!any_lane(big)) {
#pragma unroll
    for (int p = 0; p < NP; ++p) core_small<ZI, GRAD>(c[p], o[p], nozi);
  } else {                                           // some cell holds a count above 3: Stirling forms for the whole warp
#pragma unroll
    for (int p = 0; p < NP; ++p) core_big<ZI, GRAD>(c[p], o[p], nozi);
    if (any_lane(odd)) {
#pragma unroll
      for (int p = 0; p < NP; ++p) core_general<ZI, GRAD>(c[p], o[p], nozi);
    }
  }
}

template <bool ZI, bool GRAD>
PM_HD Core core_pair(F2 mu, F2 th, F2 pi, F2 x) {
  CoreState c[1]; Core o[1];
  c[0].mu = mu; c[0].th = th; c[0].pi = pi; c[0].x = x;
  core_multi<ZI, GRAD, 1>(c, o);
  return o[0];
}

struct Elem2 { F2 llk, ga, gb, gl, mu, th; };    // ga, gb, gl = d llk / d raw head outputs (mean, dispersion, dropout)

// default links of the VAE / DCA / SISUA heads: mean = softplus(ra), dispersion = softplus(rb + log(e - 1))
template <bool ZI, bool GRAD>
PM_HD Elem2 elem_pair_softplus(F2 ra, F2 rb, F2 pi, F2 x) {
  Elem2 o;
  const F2 rbs = add(rb, bc(kSoftplus1Shift));
  const Link ka = link_begin(ra), kb = link_begin(rbs);
  F2 dmu = bc(0.f), dth = bc(0.f);
  if (GRAD) {
    const F2 R = rcp(mul(ka.s, kb.s));
    const F2 rsa = mul(R, kb.s), rsb = mul(R, ka.s);
    o.mu = link_value(ka, ra, rsa);
    o.th = link_value(kb, rbs, rsb);
    dmu = mul(ka.e, rsa); dth = mul(kb.e, rsb);
  } else {
    o.mu = link_value_nograd(ka, ra);
    o.th = link_value_nograd(kb, rbs);
  }
  const Core c = core_pair<ZI, GRAD>(o.mu, o.th, pi, x);
  o.llk = c.llk;
  o.ga = mul(c.gmu, dmu); o.gb = mul(c.gth, dth); o.gl = c.gl;
  return o;
}

// scVI links (scvi.py:117-138): mean = exp(clip(library)) * clamp(softmax_g(u)), dispersion = exp(rb).
//   u_lse = u_g - logsumexp_g(u), eL = exp(clipped library).  t = d llk / d s_raw_g (softmax output before the clamp);
//   the caller finishes the softmax Jacobian once the row sum of s_raw t is known.  gmu_mu -> d llk / d library.
struct Scvi2 { F2 llk, mu, th, s_raw, t, gmu_mu, gb, gl; };
template <bool ZI, bool GRAD>
PM_HD Scvi2 elem_pair_scvi(F2 u_lse, F2 rb, F2 pi, F2 x, float eL) {
  Scvi2 o;
  const float lo = 1e-7f, hi = 1.f - 1e-7f;
  o.s_raw = ex2(mul(u_lse, bc(kLog2e)));
  o.mu = mul(bc(eL), min2(max2(o.s_raw, lo), hi));
  o.th = ex2(mul(rb, bc(kLog2e)));
  const Core c = core_pair<ZI, GRAD>(o.mu, o.th, pi, x);
  o.llk = c.llk;
  const F2 ge = mul(c.gmu, bc(eL));
  o.t = mk((o.s_raw.x >= lo && o.s_raw.x <= hi) ? ge.x : 0.f, (o.s_raw.y >= lo && o.s_raw.y <= hi) ? ge.y : 0.f);
  o.gmu_mu = mul(c.gmu, o.mu);
  o.gb = mul(c.gth, o.th);
  o.gl = c.gl;
  return o;
}

// NP pairs at once (independent dependency chains the compiler interleaves)
template <bool ZI, bool GRAD, int NP>
PM_HD void elem_multi_softplus(const F2 (&ra)[NP], const F2 (&rb)[NP], const F2 (&pi)[NP], const F2 (&x)[NP], Elem2 (&o)[NP],
                               bool nozi = false) {
  CoreState c[NP]; Core k[NP];
  F2 dmu[NP], dth[NP];
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    const F2 rbs = add(rb[p], bc(kSoftplus1Shift));
    const Link ka = link_begin(ra[p]), kb = link_begin(rbs);
    dmu[p] = bc(0.f); dth[p] = bc(0.f);
    if (GRAD) {
      const F2 R = rcp(mul(ka.s, kb.s));
      const F2 rsa = mul(R, kb.s), rsb = mul(R, ka.s);
      c[p].mu = link_value(ka, ra[p], rsa);
      c[p].th = link_value(kb, rbs, rsb);
      dmu[p] = mul(ka.e, rsa); dth[p] = mul(kb.e, rsb);
    } else {
      c[p].mu = link_value_nograd(ka, ra[p]);
      c[p].th = link_value_nograd(kb, rbs);
    }
    c[p].pi = pi[p]; c[p].x = x[p];
  }
  core_multi<ZI, GRAD, NP>(c, k, nozi);
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    o[p].mu = c[p].mu; o[p].th = c[p].th; o[p].llk = k[p].llk;
    o[p].ga = mul(k[p].gmu, dmu[p]); o[p].gb = mul(k[p].gth, dth[p]); o[p].gl = k[p].gl;
  }
}

template <bool ZI, bool GRAD, int NP>
PM_HD void elem_multi_scvi(const F2 (&u_lse)[NP], const F2 (&rb)[NP], const F2 (&pi)[NP], const F2 (&x)[NP], float eL, Scvi2 (&o)[NP],
                           bool nozi = false) {
  const float lo = 1e-7f, hi = 1.f - 1e-7f;
  CoreState c[NP]; Core k[NP];
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    o[p].s_raw = ex2(mul(u_lse[p], bc(kLog2e)));
    c[p].mu = mul(bc(eL), min2(max2(o[p].s_raw, lo), hi));
    c[p].th = ex2(mul(rb[p], bc(kLog2e)));
    c[p].pi = pi[p]; c[p].x = x[p];
  }
  core_multi<ZI, GRAD, NP>(c, k, nozi);
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    o[p].mu = c[p].mu; o[p].th = c[p].th; o[p].llk = k[p].llk;
    const F2 ge = mul(k[p].gmu, bc(eL));
    o[p].t = mk((o[p].s_raw.x >= lo && o[p].s_raw.x <= hi) ? ge.x : 0.f, (o[p].s_raw.y >= lo && o[p].s_raw.y <= hi) ? ge.y : 0.f);
    o[p].gmu_mu = mul(k[p].gmu, c[p].mu);
    o[p].gb = mul(k[p].gth, c[p].th);
    o[p].gl = k[p].gl;
  }
}

// TFP parameterisation of the 'zinb' / 'nb' output enums (tests/test_singlecell_models.py:60-80): head 0 = log total_count
// a, head 1 = logits b:  NegativeBinomial(total_count = e^a, logits = b)  ==  NB(mean = e^(a+b), inverse dispersion = e^a)
// (theta log(theta/(theta+mu)) = r log sigmoid(-b), x log(mu/(theta+mu)) = x log sigmoid(b)), so only the links differ:
//   d/da = gmu mu + gth theta,   d/db = gmu mu.
template <bool ZI, bool GRAD, int NP>
PM_HD void elem_multi_tfp(const F2 (&ra)[NP], const F2 (&rb)[NP], const F2 (&pi)[NP], const F2 (&x)[NP], Elem2 (&o)[NP],
                          bool nozi = false) {
  CoreState c[NP]; Core k[NP];
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    c[p].th = ex2(mul(min2(ra[p], 80.f), bc(kLog2e)));
    c[p].mu = ex2(mul(min2(add(ra[p], rb[p]), 80.f), bc(kLog2e)));
    c[p].pi = pi[p]; c[p].x = x[p]; c[p].eps = 1e-30f;
  }
  core_multi<ZI, GRAD, NP>(c, k, nozi);
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    o[p].mu = c[p].mu; o[p].th = c[p].th; o[p].llk = k[p].llk;
    const F2 gm = mul(k[p].gmu, c[p].mu);
    o[p].gb = gm;                                   // d llk / d logits
    o[p].ga = fma2(k[p].gth, c[p].th, gm);          // d llk / d log total_count
    o[p].gl = k[p].gl;
  }
}

}  // namespace pm
}  // namespace sisua
