"""Known-answer tests that pin the oracle (the reference ships no golden vectors for this
path)."""
import math

import numpy as np
import pytest
import scipy.special as sps
import scipy.stats as st
import torch

from oracle import step_oracle as O
from sisua_b200 import config as C
from sisua_b200 import params as PR
from tests import helpers as Hh

T = lambda a: torch.tensor(a, dtype=torch.float64)


def test_nb_matches_scipy():
  x = np.array([0., 1., 2., 7., 40., 300.])
  mu = np.array([0.3, 1.5, 2.0, 10.0, 35.0, 280.0])
  th = np.array([0.2, 1.0, 3.0, 0.7, 12.0, 5.0])
  ref = st.nbinom.logpmf(x, th, th / (th + mu))
  got = O.log_nb_disp(T(x), T(mu), T(th)).numpy()
  np.testing.assert_allclose(got, ref, rtol=1e-6, atol=1e-6)


def test_nb_tfp_matches_scipy():
  # TFP NegativeBinomial(total_count r, logits b): p(success) = sigmoid(b), mean r * exp(b)
  y = np.array([0., 1., 3., 9.])
  a = np.array([0.1, 0.5, 1.0, 2.0])
  b = np.array([-1.0, 0.0, 0.5, 1.2])
  r = np.exp(a)
  ref = st.nbinom.logpmf(y, r, 1.0 - sps.expit(b))
  got = O.log_nb_tfp(T(y), T(a), T(b)).numpy()
  np.testing.assert_allclose(got, ref, rtol=1e-9, atol=1e-9)


def test_zinb_mixture_and_normalisation():
  mu, th, pi_logit = 3.0, 1.7, -0.4
  xs = np.arange(0, 400, dtype=np.float64)
  lp = O.log_zinb_disp(T(xs), T(np.full_like(xs, mu)), T(np.full_like(xs, th)),
                       T(np.full_like(xs, pi_logit))).numpy()
  assert abs(np.exp(lp).sum() - 1.0) < 1e-6
  pi = sps.expit(pi_logit)
  nb = st.nbinom.pmf(xs, th, th / (th + mu))
  ref = np.log(np.where(xs == 0, pi + (1 - pi) * nb, (1 - pi) * nb))
  np.testing.assert_allclose(lp, ref, rtol=1e-6, atol=1e-6)


def test_real_valued_protein_nb_is_finite():
  y = T([0.53, 2.2, 9.11])
  v = O.log_nb_tfp(y, T([0.2, 0.2, 0.2]), T([0.1, -0.3, 1.0]))
  assert torch.isfinite(v).all()


def test_kl_closed_forms():
  loc, scale = T([[0.3, -1.2, 0.0]]), T([[0.5, 1.7, 1.0]])
  kl = O.kl_diag_normal_std(loc, scale)
  q = torch.distributions.Normal(loc, scale)
  p = torch.distributions.Normal(torch.zeros_like(loc), torch.ones_like(scale))
  ref = torch.distributions.kl_divergence(q, p).sum(-1)
  assert torch.allclose(kl, ref, atol=1e-12)
  assert float(O.kl_diag_normal_std(T([[0., 0.]]), T([[1., 1.]]))) == 0.0
  kl2 = O.kl_normal_normal(T([6.1]), T([0.4]), T([6.4]), T([0.08]))
  ref2 = torch.distributions.kl_divergence(torch.distributions.Normal(T([6.1]), T([0.4])),
                                           torch.distributions.Normal(T([6.4]), T([math.sqrt(0.08)])))
  assert torch.allclose(kl2, ref2, atol=1e-12)


def test_softplus1_candidates():
  z = T([0.0])
  assert abs(float(O.activation(C.ACT_SOFTPLUS1, z)) - 1.0) < 1e-12
  assert abs(float(O.activation(C.ACT_SOFTPLUS_P1, z)) - (math.log(2) + 1)) < 1e-12
  assert abs(float(O.activation(C.ACT_SOFTPLUS, z)) - math.log(2)) < 1e-12


def test_adam_single_step_by_hand():
  P = {"w": T([1.0, -2.0])}
  g = {"w": T([0.5, -0.25])}
  m = {"w": torch.zeros(2, dtype=torch.float64)}
  v = {"w": torch.zeros(2, dtype=torch.float64)}
  O.adam_update(P, g, m, v, t=1, lr=1e-3, clipnorm=100.0)
  # t=1: m = .1 g, v = .001 g^2, lr_t = lr*sqrt(.001)/.1 ; update = lr_t * m/(sqrt(v)+1e-7)
  lr_t = 1e-3 * math.sqrt(1 - 0.999) / (1 - 0.9)
  exp = np.array([1.0, -2.0]) - lr_t * (0.1 * np.array([0.5, -0.25])) / (
      np.sqrt(0.001 * np.array([0.25, 0.0625])) + 1e-7)
  np.testing.assert_allclose(P["w"].numpy(), exp, rtol=1e-12)


def test_clipnorm_per_variable():
  P = {"a": T([0.0, 0.0]), "b": T([0.0])}
  g = {"a": T([300.0, 400.0]), "b": T([1.0])}
  m = {k: torch.zeros_like(p) for k, p in P.items()}
  v = {k: torch.zeros_like(p) for k, p in P.items()}
  O.adam_update(P, g, m, v, t=1, clipnorm=100.0, clip_mode=0)
  np.testing.assert_allclose(m["a"].numpy(), 0.1 * np.array([60.0, 80.0]), rtol=1e-12)
  np.testing.assert_allclose(m["b"].numpy(), [0.1], rtol=1e-12)


@pytest.mark.parametrize("model", ["vae", "scvi", "dca", "sisua"])
def test_forward_shapes_and_fp32_fp64_agree(model):
  cfg = C.make_step_config(model, n_genes=50, n_proteins=5 if model == "sisua" else 0, n_latent=6)
  flat = Hh.randomize_norm_params(cfg, PR.init_flat_params(cfg))
  mov = PR.init_bn_moving(cfg)
  batch = Hh.make_batch(cfg, 32)
  o64 = O.forward(cfg, Hh.oracle_params(cfg, flat), Hh.oracle_moving(cfg, mov), training=True, **batch)
  o32 = O.forward(cfg, Hh.oracle_params(cfg, flat, torch.float32),
                  Hh.oracle_moving(cfg, mov, torch.float32), training=True, **batch)
  assert o64["elbo"].shape == (32,)
  assert o64["mu"].shape == (32, 50)
  assert o64["z_loc"].shape == (32, 6)
  np.testing.assert_allclose(o32["elbo"].numpy(), o64["elbo"].numpy(), rtol=2e-5)
  if model == "dca":
    assert float(o64["kl_z"].abs().sum()) == 0.0
  if model == "scvi":
    assert float(o64["kl_l"].abs().sum()) > 0.0


def test_mc_sample_axis():
  cfg = C.make_step_config("vae", n_genes=40, n_latent=4)
  flat = PR.init_flat_params(cfg)
  mov = PR.init_bn_moving(cfg)
  batch = Hh.make_batch(cfg, 8, S=3)
  o = O.forward(cfg, Hh.oracle_params(cfg, flat), Hh.oracle_moving(cfg, mov), training=False, **batch)
  assert o["mu"].shape == (3, 8, 40) and o["elbo"].shape == (3, 8)
  assert O.imputed_mean(o).shape == (8, 40)
  # sample s alone must equal the s-th slice (inference BN has no batch coupling)
  b1 = dict(batch, eps_z=batch["eps_z"][1])
  o1 = O.forward(cfg, Hh.oracle_params(cfg, flat), Hh.oracle_moving(cfg, mov), training=False, **b1)
  np.testing.assert_allclose(o1["elbo"].numpy(), o["elbo"][1].numpy(), rtol=1e-12)


def test_finite_difference_gradient():
  cfg = C.make_step_config("sisua", n_genes=12, n_proteins=3, n_latent=3)
  flat = Hh.randomize_norm_params(cfg, PR.init_flat_params(cfg))
  batch = Hh.make_batch(cfg, 6)
  P = Hh.oracle_params(cfg, flat)
  for p in P.values():
    p.requires_grad_(True)
  out = O.forward(cfg, P, None, training=True, **batch)
  out["loss"].backward()
  rng = np.random.default_rng(0)
  for name in ["enc.0.W", "enc.1.gamma", "lat.W", "dec.0.W", "out.W", "out.b", "y.W"]:
    p = P[name]
    idx = tuple(int(rng.integers(0, s)) for s in p.shape)
    h = 1e-6
    with torch.no_grad():
      old = float(p[idx]); p[idx] = old + h
      lp = float(O.forward(cfg, P, None, training=True, **batch)["loss"])
      p[idx] = old - h
      lm = float(O.forward(cfg, P, None, training=True, **batch)["loss"])
      p[idx] = old
    fd = (lp - lm) / (2 * h)
    assert abs(fd - float(p.grad[idx])) <= 1e-5 * max(1.0, abs(fd)), (name, fd, float(p.grad[idx]))


def test_library_stats_follow_reference_recipe():
  X = np.array([[1., 2., 3.], [0., 0., 10.], [5., 5., 5.]], dtype=np.float32)
  lc = np.log(X.sum(1) + 1e-8)
  s = O.library_size_stats(X)
  np.testing.assert_allclose(s[:, 0], lc.mean(), rtol=1e-6)
  np.testing.assert_allclose(s[:, 1], lc.var(), rtol=1e-6)


def test_philox_known_answers():
  # Random123 kat_vectors for philox4x32-10
  from oracle.philox import dropout_mask, philox4x32_10
  kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
         ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
         ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
          (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
  for ctr, key, exp in kat:
    got = tuple(int(v) for v in philox4x32_10(*ctr, *key))
    assert got == exp
  m = dropout_mask(512, 200, 0.3, seed=1234, step=7, stream=0)
  assert m.shape == (512, 200) and abs(m.mean() - 0.7) < 0.01
  assert not np.array_equal(m, dropout_mask(512, 200, 0.3, seed=1234, step=8, stream=0))
