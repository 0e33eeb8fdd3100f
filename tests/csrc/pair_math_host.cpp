// Host build of sisua_b200/csrc/pair_math.cuh (plain float arithmetic stands in for the packed / MUFU instructions) so
// the formulas of the fused epilogue can be checked against a float64 restatement without a GPU
// (tests/test_pair_math_cpu.py builds this with g++ and calls it through ctypes).  TEST INFRASTRUCTURE ONLY.
#include <cmath>
#include <cstdint>

namespace sisua { namespace pm {
static double digamma(double t) {
  double r = 0.0;
  while (t < 10.0) { r -= 1.0 / t; t += 1.0; }
  const double it = 1.0 / t, it2 = it * it;
  return r + std::log(t) - 0.5 * it - it2 * (1.0 / 12 - it2 * (1.0 / 120 - it2 * (1.0 / 252)));
}
// scalar fall-back of the device build = device_math.cuh:count_core_fast<.., 1>; here: the same closed forms in double
template <bool ZI, bool GRAD>
inline void core_scalar_fallback(float muf, float thf, float pif, float xf, float epsf, float& llk, float& gmu, float& gth, float& gl) {
  const double mu = muf, th = thf, pi = pif, x = xf, eps = epsf;
  const double ltm = std::log(th + mu + eps), dlog = std::log(th + eps) - ltm, n0 = th * dlog;
  const double rt = 1.0 / (th + mu + eps);
  const double dn0_dmu = -th * rt, dn0_dth = dlog + th * (1.0 / (th + eps) - rt);
  auto sp = [](double v) { return v > 0 ? v + std::log1p(std::exp(-v)) : std::log1p(std::exp(v)); };
  auto sg = [](double v) { return 1.0 / (1.0 + std::exp(-v)); };
  if (x < 1e-8) {
    if (ZI) { llk = (float)(sp(n0 - pi) - sp(-pi)); const double w = sg(n0 - pi); gl = (float)(sg(-pi) - w); gmu = (float)(w * dn0_dmu); gth = (float)(w * dn0_dth); }
    else { llk = (float)n0; gmu = (float)dn0_dmu; gth = (float)dn0_dth; gl = 0.f; }
  } else {
    double l = n0 + x * (std::log(mu + eps) - ltm) + std::lgamma(x + th) - std::lgamma(th) - std::lgamma(x + 1.0);
    double gm = dn0_dmu + x * (1.0 / (mu + eps) - rt), gt = dn0_dth - x * rt + digamma(x + th) - digamma(th), gg = 0.0;
    if (ZI) { l -= sp(pi); gg = -sg(pi); }
    llk = (float)l; gmu = (float)gm; gth = (float)gt; gl = (float)gg;
  }
}
}}  // namespace

#include "../../sisua_b200/csrc/pair_math.cuh"

using namespace sisua::pm;

// out[n][6] = llk, ga, gb, gl, mu, th for n elements evaluated as n/2 pairs (n even)
extern "C" void pm_elem_softplus(const float* ra, const float* rb, const float* pi, const float* x, int n, int zi, int grad, float* out) {
  for (int i = 0; i + 1 < n; i += 2) {
    Elem2 e;
    const F2 a = mk(ra[i], ra[i + 1]), b = mk(rb[i], rb[i + 1]), p = mk(pi[i], pi[i + 1]), c = mk(x[i], x[i + 1]);
    if (zi) e = grad ? elem_pair_softplus<true, true>(a, b, p, c) : elem_pair_softplus<true, false>(a, b, p, c);
    else e = grad ? elem_pair_softplus<false, true>(a, b, p, c) : elem_pair_softplus<false, false>(a, b, p, c);
    float* o = out + (size_t)i * 6;
    o[0] = e.llk.x; o[1] = e.ga.x; o[2] = e.gb.x; o[3] = e.gl.x; o[4] = e.mu.x; o[5] = e.th.x;
    o[6] = e.llk.y; o[7] = e.ga.y; o[8] = e.gb.y; o[9] = e.gl.y; o[10] = e.mu.y; o[11] = e.th.y;
  }
}

// out[n][8] = llk, mu, th, s_raw, t, gmu_mu, gb, gl
extern "C" void pm_elem_scvi(const float* u_lse, const float* rb, const float* pi, const float* x, const float* eL, int n, int zi,
                             int grad, float* out) {
  for (int i = 0; i + 1 < n; i += 2) {
    Scvi2 e;
    const F2 a = mk(u_lse[i], u_lse[i + 1]), b = mk(rb[i], rb[i + 1]), p = mk(pi[i], pi[i + 1]), c = mk(x[i], x[i + 1]);
    if (zi) e = grad ? elem_pair_scvi<true, true>(a, b, p, c, eL[i]) : elem_pair_scvi<true, false>(a, b, p, c, eL[i]);
    else e = grad ? elem_pair_scvi<false, true>(a, b, p, c, eL[i]) : elem_pair_scvi<false, false>(a, b, p, c, eL[i]);
    float* o = out + (size_t)i * 8;
    o[0] = e.llk.x; o[1] = e.mu.x; o[2] = e.th.x; o[3] = e.s_raw.x; o[4] = e.t.x; o[5] = e.gmu_mu.x; o[6] = e.gb.x; o[7] = e.gl.x;
    o[8] = e.llk.y; o[9] = e.mu.y; o[10] = e.th.y; o[11] = e.s_raw.y; o[12] = e.t.y; o[13] = e.gmu_mu.y; o[14] = e.gb.y; o[15] = e.gl.y;
  }
}

// zero-inflated head evaluated without its zero inflation ("imputed" distribution): out[n][3] = llk, mu, th
extern "C" void pm_elem_softplus_nozi(const float* ra, const float* rb, const float* pi, const float* x, int n, float* out) {
  for (int i = 0; i + 1 < n; i += 2) {
    F2 a[1] = {mk(ra[i], ra[i + 1])}, b[1] = {mk(rb[i], rb[i + 1])}, p[1] = {mk(pi[i], pi[i + 1])}, c[1] = {mk(x[i], x[i + 1])};
    Elem2 e[1];
    elem_multi_softplus<true, false, 1>(a, b, p, c, e, true);
    out[i * 3] = e[0].llk.x; out[i * 3 + 1] = e[0].mu.x; out[i * 3 + 2] = e[0].th.x;
    out[i * 3 + 3] = e[0].llk.y; out[i * 3 + 4] = e[0].mu.y; out[i * 3 + 5] = e[0].th.y;
  }
}

// TFP links ('zinb' / 'nb'): out[n][6] = llk, ga (d / d log total_count), gb (d / d logits), gl, mean, total_count
extern "C" void pm_elem_tfp(const float* ra, const float* rb, const float* pi, const float* x, int n, int zi, float* out) {
  for (int i = 0; i + 1 < n; i += 2) {
    F2 a[1] = {mk(ra[i], ra[i + 1])}, b[1] = {mk(rb[i], rb[i + 1])}, p[1] = {mk(pi[i], pi[i + 1])}, c[1] = {mk(x[i], x[i + 1])};
    Elem2 e[1];
    if (zi) elem_multi_tfp<true, true, 1>(a, b, p, c, e); else elem_multi_tfp<false, true, 1>(a, b, p, c, e);
    float* o = out + (size_t)i * 6;
    o[0] = e[0].llk.x; o[1] = e[0].ga.x; o[2] = e[0].gb.x; o[3] = e[0].gl.x; o[4] = e[0].mu.x; o[5] = e[0].th.x;
    o[6] = e[0].llk.y; o[7] = e[0].ga.y; o[8] = e[0].gb.y; o[9] = e[0].gl.y; o[10] = e[0].mu.y; o[11] = e[0].th.y;
  }
}

extern "C" void pm_ex2_poly(const float* t, int n, float* out) {
  for (int i = 0; i + 1 < n; i += 2) { const F2 r = ex2_poly(mk(t[i], t[i + 1])); out[i] = r.x; out[i + 1] = r.y; }
}
