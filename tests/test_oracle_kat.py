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


def test_oracle_matches_torch_nn_composition():
  """The oracle's ZINB-VAE train-mode forward against an independent composition out of stock building blocks
  (nn.Linear, nn.BatchNorm1d(eps=1e-3), torch.distributions.Normal / NegativeBinomial / kl_divergence): per-cell ELBO,
  its terms, the updated moving statistics and the gradient of the loss wrt the first weight matrix."""
  import torch.nn as nn
  import torch.distributions as TD
  cfg = C.make_step_config("vae", n_genes=30, n_latent=5, input_dropout=0.0)
  flat = Hh.randomize_norm_params(cfg, PR.init_flat_params(cfg))
  mov = PR.init_bn_moving(cfg)
  batch = Hh.make_batch(cfg, 24)
  P = Hh.oracle_params(cfg, flat)
  for p in P.values():
    p.requires_grad_(True)
  o = O.forward(cfg, P, Hh.oracle_moving(cfg, mov), training=True, **batch)
  (-o["elbo"].mean()).backward()

  dt = torch.float64
  W = {k: v.detach().clone() for k, v in P.items()}
  def unit(prefix, n_in):
    lin = nn.Linear(n_in, 64, bias=False).to(dt)
    bn = nn.BatchNorm1d(64, eps=cfg.bn_eps, momentum=1.0 - cfg.bn_momentum).to(dt)
    with torch.no_grad():
      lin.weight.copy_(W[prefix + ".W"][:, :n_in]); bn.weight.copy_(W[prefix + ".gamma"]); bn.bias.copy_(W[prefix + ".beta"])
    return lin, bn
  x = torch.tensor(batch["x"], dtype=dt)
  h = torch.log1p(x)
  mods = []
  first = None
  for i, (name, n_in) in enumerate([("enc.0", 30), ("enc.1", 64)]):
    lin, bn = unit(name, n_in)
    mods.append((name, bn))
    if first is None:
      first = lin
    h = torch.relu(bn.train()(lin(h)))
  pl = h @ W["lat.W"].T + W["lat.b"]
  loc, scale = pl[:, :5], torch.nn.functional.softplus(pl[:, 5:] + np.log(np.e - 1.0))
  qz = TD.Normal(loc, scale)
  z = loc + scale * torch.tensor(batch["eps_z"], dtype=dt)
  kl = TD.kl_divergence(qz, TD.Normal(torch.zeros_like(loc), torch.ones_like(scale))).sum(-1)
  d = z
  for name, n_in in [("dec.0", 5), ("dec.1", 64)]:
    lin, bn = unit(name, n_in)
    mods.append((name, bn))
    d = torch.relu(bn.train()(lin(d)))
  G = 30
  out = d @ W["out.W"].T + W["out.b"]
  mu = torch.nn.functional.softplus(out[:, :G])
  theta = torch.nn.functional.softplus(out[:, G:2 * G] + np.log(np.e - 1.0))
  pi_logit = out[:, 2 * G:]
  # NB(mean, inverse dispersion) == NegativeBinomial(total_count = theta, probs = mu / (mu + theta))
  nb = TD.NegativeBinomial(total_count=theta, probs=mu / (mu + theta))
  log_nb = nb.log_prob(x)
  log_pi, log_1m_pi = torch.nn.functional.logsigmoid(pi_logit), torch.nn.functional.logsigmoid(-pi_logit)
  llk = torch.where(x < 1e-8, torch.logsumexp(torch.stack([log_pi, log_1m_pi + log_nb]), 0), log_1m_pi + log_nb).sum(-1)
  elbo = llk - kl
  np.testing.assert_allclose(o["llk_x"].detach().numpy(), llk.detach().numpy(), rtol=1e-5)
  np.testing.assert_allclose(o["kl_z"].detach().numpy(), kl.detach().numpy(), rtol=1e-9)
  np.testing.assert_allclose(o["elbo"].detach().numpy(), elbo.detach().numpy(), rtol=1e-5)
  (-elbo.mean()).backward()
  g_ref = first.weight.grad.numpy()
  g_or = P["enc.0.W"].grad.numpy()[:, :30]
  np.testing.assert_allclose(g_or, g_ref, rtol=1e-4, atol=1e-9 + 1e-6 * np.abs(g_ref).max())
  # moving statistics (Keras: moving <- m * moving + (1 - m) * batch, biased batch variance; torch keeps the unbiased one)
  B = 24
  for name, bn in mods:
    np.testing.assert_allclose(o["new_moving"][name + ".mean"].numpy(), bn.running_mean.numpy(), rtol=1e-9, atol=1e-12)
    keras_var = (bn.running_var.numpy() - cfg.bn_momentum * 1.0) * (B - 1) / B + cfg.bn_momentum * 1.0
    np.testing.assert_allclose(o["new_moving"][name + ".var"].numpy(), keras_var, rtol=1e-8)


def test_oracle_scvi_and_sisua_heads_match_torch_distributions():
  """Inference-mode (moving-statistics BatchNorm) forward of scVI and SISUA against stock torch modules: library
  latent with its dataset prior, gene softmax scaled by the clipped library, exp dispersion, and the protein head as
  TFP's NegativeBinomial(total_count = exp(a), logits = b), masked and weighted by alpha."""
  import torch.nn.functional as F
  import torch.distributions as TD
  dt = torch.float64

  def hidden(P, M, prefix, n_layers, h):
    for i in range(n_layers):
      n = f"{prefix}.{i}"
      a = h @ P[n + ".W"][:, :h.shape[1]].T
      a = (a - M[n + ".mean"]) / torch.sqrt(M[n + ".var"] + 1e-3) * P[n + ".gamma"] + P[n + ".beta"]
      h = torch.relu(a)
    return h

  def zinb(x, mu, theta, pi_logit):
    log_nb = TD.NegativeBinomial(total_count=theta, probs=mu / (mu + theta)).log_prob(x)
    lp, l1 = F.logsigmoid(pi_logit), F.logsigmoid(-pi_logit)
    return torch.where(x < 1e-8, torch.logsumexp(torch.stack([lp, l1 + log_nb]), 0), l1 + log_nb).sum(-1)

  for model, kw in (("scvi", dict(clip_library=7.0)), ("sisua", dict(n_proteins=4, alpha=10.0))):
    cfg = C.make_step_config(model, n_genes=20, n_latent=3, **kw)
    flat = Hh.randomize_norm_params(cfg, PR.init_flat_params(cfg))
    mov = PR.init_bn_moving(cfg)
    rng = np.random.default_rng(3)
    mov[:, 0, :] = rng.normal(0, 0.2, mov[:, 0, :].shape); mov[:, 1, :] = rng.uniform(0.5, 1.5, mov[:, 1, :].shape)
    batch = Hh.make_batch(cfg, 16)
    P, M = Hh.oracle_params(cfg, flat), Hh.oracle_moving(cfg, mov)
    o = O.forward(cfg, P, M, training=False, **batch)
    x = torch.tensor(batch["x"], dtype=dt)
    G, Z = 20, 3
    h = hidden(P, M, "enc", cfg.n_enc_layers, torch.log1p(x))
    pl = h @ P["lat.W"].T + P["lat.b"]
    loc, scale = pl[:, :Z], F.softplus(pl[:, Z:] + np.log(np.e - 1.0))
    z = loc + scale * torch.tensor(batch["eps_z"], dtype=dt)
    kl = TD.kl_divergence(TD.Normal(loc, scale), TD.Normal(torch.zeros_like(loc), torch.ones_like(scale))).sum(-1)
    d = hidden(P, M, "dec", cfg.n_dec_layers, z)
    out = d @ P["out.W"].T + P["out.b"]
    if model == "scvi":
      hl = hidden(P, M, "encl", cfg.n_encl_layers, torch.log1p(x))
      pll = hl @ P["lib.W"].T + P["lib.b"]
      l_loc, l_scale = pll[:, 0], F.softplus(pll[:, 1] + np.log(np.e - 1.0))
      lib = l_loc + l_scale * torch.tensor(batch["eps_l"], dtype=dt)
      prior = torch.tensor(batch["library"], dtype=dt)
      kl_l = TD.kl_divergence(TD.Normal(l_loc, l_scale), TD.Normal(prior[:, 0], torch.sqrt(prior[:, 1])))
      mu = torch.exp(lib.clamp(0.0, 7.0))[:, None] * torch.softmax(out[:, :G], 1).clamp(1e-7, 1 - 1e-7)
      llk = zinb(x, mu, torch.exp(out[:, G:2 * G]), out[:, 2 * G:])
      elbo = llk - kl - kl_l
      np.testing.assert_allclose(o["kl_l"].numpy(), kl_l.numpy(), rtol=1e-9)
      assert float((lib > 7.0).sum()) > 0 or float(lib.max()) <= 7.0      # clip path exercised or trivially inactive
    else:
      mu = F.softplus(out[:, :G])
      llk = zinb(x, mu, F.softplus(out[:, G:2 * G] + np.log(np.e - 1.0)), out[:, 2 * G:])
      py = d @ P["y.W"].T + P["y.b"]
      y = torch.tensor(batch["y"], dtype=dt)
      llk_y = TD.NegativeBinomial(total_count=torch.exp(py[:, :4]), logits=py[:, 4:], validate_args=False).log_prob(y).sum(-1)   # y is real-valued
      m = torch.tensor(batch["mask"], dtype=dt)
      elbo = llk + 10.0 * m * llk_y - kl
      np.testing.assert_allclose(o["llk_y"].numpy(), llk_y.numpy(), rtol=1e-6)
    np.testing.assert_allclose(o["mu"].numpy(), mu.numpy(), rtol=1e-9)
    np.testing.assert_allclose(o["llk_x"].numpy(), llk.numpy(), rtol=1e-5)
    np.testing.assert_allclose(o["elbo"].numpy(), elbo.numpy(), rtol=1e-5)


def test_philox_normal_noise_is_standard_normal_and_counter_based():
  """The reparameterisation noise generator (Box-Muller on Philox words): distribution, independence across columns /
  steps / streams, and pure-function behaviour (what lets the backward pass regenerate it)."""
  from scipy import stats
  from oracle.philox import NOISE_STREAM_L, NOISE_STREAM_Z, normal_noise
  n = normal_noise(40000, 10, 0xDEADBEEFCAFE, 3, NOISE_STREAM_Z)
  assert n.shape == (40000, 10)
  assert stats.kstest(n.ravel(), "norm").pvalue > 1e-3
  assert abs(n.mean()) < 5e-3 and abs(n.std() - 1) < 5e-3
  c = np.corrcoef(n.T)
  assert np.abs(c - np.eye(10)).max() < 0.02
  assert np.array_equal(n, normal_noise(40000, 10, 0xDEADBEEFCAFE, 3, NOISE_STREAM_Z))
  assert np.array_equal(n[:100, :6], normal_noise(100, 6, 0xDEADBEEFCAFE, 3, NOISE_STREAM_Z))     # ragged widths share the stream
  for other in (normal_noise(40000, 10, 0xDEADBEEFCAFE, 4, NOISE_STREAM_Z), normal_noise(40000, 10, 0xDEADBEEFCAFE, 3, NOISE_STREAM_L),
                normal_noise(40000, 10, 0xDEADBEEFCAFF, 3, NOISE_STREAM_Z)):
    assert abs(np.corrcoef(n.ravel(), other.ravel())[0, 1]) < 0.01


def test_gpu_corruption_restatement_statistics():
  """oracle/philox.py:corrupt_counts (the bit-exact mirror of sisua_corrupt_counts): zeros stay zero, nothing grows, the
  selected share and the retained mass follow dropout / retain_rate, 'uniform' only zeroes, a seed reproduces."""
  from oracle import philox as PH
  rng = np.random.default_rng(0)
  x = (rng.poisson(0.6, (400, 300)) * (rng.random((400, 300)) < 0.7)).astype(np.float32)
  x[0, 0] = 700.0
  nz = x > 0
  y = PH.corrupt_counts(x, 0.25, 0.2, "binomial", seed=11)
  assert (y[~nz] == 0).all() and (y <= x).all() and (y == np.floor(y)).all()
  assert abs(y[nz].sum() / x[nz].sum() - (1 - 0.25 * 0.8)) < 0.02
  np.testing.assert_array_equal(y, PH.corrupt_counts(x, 0.25, 0.2, "binomial", seed=11))
  assert (y != PH.corrupt_counts(x, 0.25, 0.2, "binomial", seed=12)).any()
  u = PH.corrupt_counts(x, 0.25, 0.2, "uniform", seed=11)
  assert ((u == x) | (u == 0)).all() and abs((u[nz] > 0).mean() - (1 - 0.25 * 0.8)) < 0.02
  np.testing.assert_array_equal(PH.corrupt_counts(x, 0.0, 0.2, "binomial", seed=3), x)       # nothing selected
  full = PH.corrupt_counts(x, 0.999, 1.0, "binomial", seed=3)                                   # everything retained
  np.testing.assert_array_equal(full, x)
