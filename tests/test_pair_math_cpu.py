"""Numerics of the pair-wise epilogue mathematics (sisua_b200/csrc/pair_math.cuh) without a GPU: the header is built
for the host with g++ (plain float arithmetic in place of the packed / MUFU instructions; tests/csrc/pair_math_host.cpp)
and compared with the float64 oracle formulas and their autograd derivatives over a grid that covers saturated links,
clamped dropout logits, zero / small / large / non-integer counts."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import step_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIFT = 0.5413248546129181


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
  out = str(tmp_path_factory.mktemp("pm") / "libpm_host.so")
  src = os.path.join(ROOT, "tests", "csrc", "pair_math_host.cpp")
  r = subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out, src], capture_output=True, text=True)
  assert r.returncode == 0, r.stderr
  L = ctypes.CDLL(out)
  fp = ctypes.POINTER(ctypes.c_float)
  L.pm_elem_softplus.argtypes = [fp, fp, fp, fp, ctypes.c_int, ctypes.c_int, ctypes.c_int, fp]
  L.pm_elem_scvi.argtypes = [fp, fp, fp, fp, fp, ctypes.c_int, ctypes.c_int, ctypes.c_int, fp]
  L.pm_ex2_poly.argtypes = [fp, ctypes.c_int, fp]
  L.pm_elem_tfp.argtypes = [fp, fp, fp, fp, ctypes.c_int, ctypes.c_int, fp]
  L.pm_elem_softplus_nozi.argtypes = [fp, fp, fp, fp, ctypes.c_int, fp]
  return L


def _p(a):
  return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _grid(seed=0, n=40000):
  rng = np.random.default_rng(seed)
  ra = rng.uniform(-14, 8, n); rb = rng.uniform(-10, 8, n); pi = rng.uniform(-12, 12, n)
  # saturated links / logits
  k = n // 20
  ra[:k] = rng.uniform(30, 60, k); rb[k:2 * k] = rng.uniform(30, 55, k); pi[2 * k:3 * k] = rng.uniform(35, 70, k)
  pi[3 * k:4 * k] = rng.uniform(-70, -35, k); ra[4 * k:5 * k] = rng.uniform(-40, -15, k); rb[5 * k:6 * k] = rng.uniform(-25, -10, k)
  x = rng.choice([0, 0, 0, 0, 1, 1, 2, 3, 4, 5, 8, 9, 20, 150], size=n).astype(np.float64)
  x[::97] = 2.5                      # non-integer counts take the general path
  # pairs where one lane is zero and the other large, both small, ...
  return [v.astype(np.float32) for v in (ra, rb, pi, x)]


def _reference(ra, rb, pi, x, zi):
  a = torch.tensor(ra, dtype=torch.float64, requires_grad=True)
  b = torch.tensor(rb, dtype=torch.float64, requires_grad=True)
  p = torch.tensor(pi, dtype=torch.float64, requires_grad=True)
  xx = torch.tensor(x, dtype=torch.float64)
  mu = F.softplus(a); th = F.softplus(b + SHIFT)
  llk = O.log_zinb_disp(xx, mu, th, p) if zi else O.log_nb_disp(xx, mu, th)
  llk.sum().backward()
  z = np.zeros_like(ra, dtype=np.float64)
  return (llk.detach().numpy(), a.grad.numpy(), b.grad.numpy(), p.grad.numpy() if zi else z, mu.detach().numpy(), th.detach().numpy())


@pytest.mark.parametrize("zi", [1, 0])
@pytest.mark.parametrize("grad", [1, 0])
def test_softplus_pair_matches_float64(lib, zi, grad):
  ra, rb, pi, x = _grid()
  n = ra.size
  out = np.zeros((n, 6), dtype=np.float32)
  lib.pm_elem_softplus(_p(ra), _p(rb), _p(pi), _p(x), n, zi, grad, _p(out))
  llk, ga, gb, gl, mu, th = _reference(ra, rb, pi, x, zi)
  assert np.isfinite(out).all()
  # parameters: 1e-6 relative (what the 1e-4 bound on imputed means needs, with margin)
  np.testing.assert_allclose(out[:, 4], mu, rtol=2e-6, atol=1e-30)
  np.testing.assert_allclose(out[:, 5], th, rtol=2e-6, atol=1e-30)
  # per-entry log-likelihood: a cell's llk is a sum of ~2000 of these and must hold 1e-4 relative
  err = np.abs(out[:, 0] - llk)
  assert (err <= 2e-5 * np.abs(llk) + 5e-6).all(), f"llk worst {err.max():.3e} at {np.argmax(err)}"
  if grad:
    for got, ref, name in ((out[:, 1], ga, "ga"), (out[:, 2], gb, "gb"), (out[:, 3], gl, "gl")):
      e = np.abs(got - ref)
      assert (e <= 2e-5 * np.abs(ref) + 2e-6 * (1.0 + np.abs(x))).all(), f"{name} worst {e.max():.3e} at {np.argmax(e)}"


def test_zero_inflated_head_without_zero_inflation(lib):
  """`nozi`: the NB part of a zero-inflated head (what the reference calls the "imputed" distribution,
  sisua/analysis/posterior.py:210-220) = the plain NB log-likelihood of the same mean / dispersion."""
  ra, rb, pi, x = _grid(seed=5, n=20000)
  out = np.zeros((ra.size, 3), dtype=np.float32)
  lib.pm_elem_softplus_nozi(_p(ra), _p(rb), _p(pi), _p(x), ra.size, _p(out))
  llk = _reference(ra, rb, pi, x, 0)[0]
  err = np.abs(out[:, 0] - llk)
  assert (err <= 2e-5 * np.abs(llk) + 5e-6).all(), err.max()


@pytest.mark.parametrize("zi", [1, 0])
def test_tfp_links_match_float64(lib, zi):
  """'zinb' / 'nb' output enums: TFP NegativeBinomial(total_count = e^a, logits = b) (+ zero inflation) against the
  oracle's log_nb_tfp and its autograd derivatives."""
  rng = np.random.default_rng(7)
  n = 20000
  ra = rng.uniform(-4, 5, n).astype(np.float32); rb = rng.uniform(-8, 4, n).astype(np.float32)
  pi = rng.uniform(-10, 10, n).astype(np.float32)
  x = rng.choice([0, 0, 0, 1, 2, 3, 5, 9, 40], size=n).astype(np.float32)
  out = np.zeros((n, 6), dtype=np.float32)
  lib.pm_elem_tfp(_p(ra), _p(rb), _p(pi), _p(x), n, zi, _p(out))
  a = torch.tensor(ra, dtype=torch.float64, requires_grad=True); b = torch.tensor(rb, dtype=torch.float64, requires_grad=True)
  p = torch.tensor(pi, dtype=torch.float64, requires_grad=True); xx = torch.tensor(x, dtype=torch.float64)
  base = O.log_nb_tfp(xx, a, b)
  if zi:
    llk = torch.where(xx < 1e-8, F.softplus(base - p) - F.softplus(-p), base - F.softplus(p))
  else:
    llk = base
  llk.sum().backward()
  th = np.exp(ra.astype(np.float64))
  for got, ref, name in ((out[:, 0], llk.detach().numpy(), "llk"), (out[:, 1], a.grad.numpy(), "ga"), (out[:, 2], b.grad.numpy(), "gb")) + \
      (((out[:, 3], p.grad.numpy(), "gl"),) if zi else ()):
    e = np.abs(got - ref)
    assert (e <= 3e-5 * np.abs(ref) + 5e-6 * (1 + np.abs(x)) + 3e-7 * th).all(), f"{name} worst {e.max():.3e} at {np.argmax(e)}"
  np.testing.assert_allclose(out[:, 4], np.exp(ra.astype(np.float64) + rb), rtol=3e-6)


def test_exp2_polynomial(lib):
  t = np.concatenate([np.linspace(-126, 60, 200001), np.linspace(-1, 1, 20001)]).astype(np.float32)
  t = t[: t.size // 2 * 2].copy()
  out = np.zeros_like(t)
  lib.pm_ex2_poly(_p(t), t.size, _p(out))
  ref = np.exp2(t.astype(np.float64))
  ok = t >= -125
  assert np.max(np.abs(out[ok] / ref[ok] - 1)) < 4e-7


@pytest.mark.parametrize("zi", [1, 0])
def test_scvi_pair_matches_float64(lib, zi):
  rng = np.random.default_rng(3)
  n = 20000
  u = rng.uniform(-20, -0.001, n).astype(np.float32); u[:50] = -1e-9; u[50:100] = -17.0      # clamp edges of the softmax output
  rb = rng.uniform(-6, 6, n).astype(np.float32); pi = rng.uniform(-10, 10, n).astype(np.float32)
  x = rng.choice([0, 0, 0, 1, 2, 3, 6, 30], size=n).astype(np.float32)
  eL = np.repeat(rng.uniform(50, 3000, n // 2), 2).astype(np.float32)
  out = np.zeros((n, 8), dtype=np.float32)
  lib.pm_elem_scvi(_p(u), _p(rb), _p(pi), _p(x), _p(eL), n, zi, 1, _p(out))
  s = torch.tensor(np.exp(u.astype(np.float64)), requires_grad=True)
  b = torch.tensor(rb, dtype=torch.float64, requires_grad=True)
  p = torch.tensor(pi, dtype=torch.float64, requires_grad=True)
  mu = torch.tensor(eL, dtype=torch.float64) * torch.clamp(s, 1e-7, 1 - 1e-7)
  mu.retain_grad()
  th = torch.exp(b)
  xx = torch.tensor(x, dtype=torch.float64)
  llk = O.log_zinb_disp(xx, mu, th, p) if zi else O.log_nb_disp(xx, mu, th)
  llk.sum().backward()
  e = np.abs(out[:, 0] - llk.detach().numpy())
  # theta * log(theta / (theta + mu)) in float32 carries an absolute error of ~1e-7 * theta whatever the formulation
  # (the reference's own float32 form theta * (log(theta) - log(theta + mu)) is an order of magnitude worse)
  assert (e <= 3e-5 * np.abs(llk.detach().numpy()) + 3e-6 + 2e-7 * th.detach().numpy()).all(), e.max()
  np.testing.assert_allclose(out[:, 1], mu.detach().numpy(), rtol=3e-6)
  np.testing.assert_allclose(out[:, 2], th.detach().numpy(), rtol=3e-6)
  # the clamp edges are decided in float32 on s_raw = exp(u): skip entries within rounding of the edges
  inner = (np.exp(u.astype(np.float64)) > 1.01e-7) & (np.exp(u.astype(np.float64)) < 1 - 1.2e-7)
  tol = lambda ref: 3e-5 * np.abs(ref) + 3e-6 * (1 + np.abs(x))
  t_ref = s.grad.numpy()
  assert (np.abs(out[:, 4] - t_ref)[inner] <= (3e-5 * np.abs(t_ref) + 1e-5 * eL * (1 + np.abs(x)))[inner]).all()
  gmu_mu = (mu.grad * mu).detach().numpy()
  assert (np.abs(out[:, 5] - gmu_mu) <= tol(gmu_mu) + 3e-5 * np.abs(x)).all()
  # d llk / d log(theta) = theta * (log(rho) + 1 - rho + ...): the bracket cancels to ~(mu / theta)^2 for theta >> mu and
  # keeps the 1e-7 absolute rounding of rho, scaled by theta
  assert (np.abs(out[:, 6] - b.grad.numpy()) <= tol(b.grad.numpy()) + 3e-7 * th.detach().numpy()).all()
  if zi:
    assert (np.abs(out[:, 7] - p.grad.numpy()) <= tol(p.grad.numpy()) + 3e-7 * th.detach().numpy()).all()


def test_large_counts_and_dispersions_stay_on_the_packed_path(lib):
  """Counts up to 5 000, means up to 3 000, inverse dispersions up to 10^4, fractional counts: all of them go through
  the packed Stirling forms (core_big) -- no scalar fall-back -- and must agree with float64.  The absolute floors
  cover theta >> mu, where theta * log(theta / (theta + mu)) cancels in fp32 whatever the formula."""
  rng = np.random.default_rng(5); n = 40000
  ra = rng.uniform(-6, 9, n); ra[:n // 4] = rng.uniform(20, 3000, n // 4)
  rb = rng.uniform(-6, 9, n); rb[n // 8:n // 2] = 10 ** rng.uniform(1, 4, n // 2 - n // 8)
  pi = rng.uniform(-8, 8, n)
  x = np.floor(10 ** rng.uniform(0, 3.7, n)); x[::7] = rng.integers(0, 12, len(x[::7])); x[::13] = rng.uniform(0.1, 30, len(x[::13]))
  ra, rb, pi, x = [v.astype(np.float32) for v in (ra, rb, pi, x)]
  out = np.zeros((n, 6), dtype=np.float32)
  lib.pm_elem_softplus(_p(ra), _p(rb), _p(pi), _p(x), n, 1, 1, _p(out))
  llk, ga, gb, gl, mu, th = _reference(ra, rb, pi, x, 1)
  assert np.isfinite(out).all()
  err = np.abs(out[:, 0] - llk)
  # (terms of size (x + theta) log(x + theta) cancel in the log-likelihood: fp32 leaves ~1e-7 of THEIR size)
  assert (err <= 3e-5 * np.abs(llk) + 3e-4 + 2.5e-6 * (x + th)).all(), f"llk worst {err.max():.3e} at {np.argmax(err)}"
  for k, g in enumerate((ga, gb, gl)):
    e = np.abs(out[:, 1 + k] - g)
    assert (e <= 2e-3 * np.abs(g) + 1e-4).all(), f"gradient {k}: worst {e.max():.3e} at {np.argmax(e)} (ref {g[np.argmax(e)]:.3e})"
