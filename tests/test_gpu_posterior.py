"""Posterior fast paths behind the C ABI against the float64 oracle: streamed predict,
log-likelihood of a SECOND count matrix under the reconstructed / imputed distributions, Monte-Carlo mean of the NB
mean, importance-weighted marginal log-likelihood, decoder-only and training-mode forward entry points."""
import numpy as np
import pytest
import torch

from oracle import step_oracle as O
from oracle.philox import NOISE_STREAM_L, NOISE_STREAM_Z, normal_noise
from sisua_b200 import config as C
from sisua_b200 import distributions as D
from sisua_b200 import params as PR
from sisua_b200 import synthetic as SY
from tests import helpers as Hh

pytestmark = pytest.mark.gpu


def _close(a, b, rtol=1e-4, atol=0.0, what=""):
  a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
  err = np.abs(a - b)
  bad = err > rtol * np.abs(b) + atol
  assert not bad.any(), f"{what}: {bad.sum()} / {bad.size} off; worst rel {np.max(err / (np.abs(b) + 1e-30)):.3e} abs {err.max():.3e}"


def _setup(model, kw, G, B, **cfgkw):
  cfg = C.make_step_config(model, n_genes=G, max_batch=max(4 * B, 512), **kw, **cfgkw)
  flat = Hh.randomize_norm_params(cfg, PR.init_flat_params(cfg))
  mov = PR.init_bn_moving(cfg)
  rng = np.random.default_rng(5)
  mov[:, 0, :] = rng.normal(0, 0.3, mov[:, 0, :].shape)
  mov[:, 1, :] = rng.uniform(0.5, 2.0, mov[:, 1, :].shape)
  return cfg, flat, mov, Hh.make_batch(cfg, B, seed=2)


def _noise(cfg, B, S, seed, step):
  eps = {"eps_z": np.stack([normal_noise(B, cfg.n_latent, seed, step, NOISE_STREAM_Z + 2 * s) for s in range(S)]).astype(np.float32)}
  if cfg.model_kind == C.MODEL_SCVI:
    eps["eps_l"] = np.stack([normal_noise(B, 1, seed, step, NOISE_STREAM_L + 2 * s)[:, 0] for s in range(S)]).astype(np.float32)
  return eps


@pytest.mark.parametrize("model,kw", [("vae", {}), ("scvi", {}), ("sisua", dict(n_proteins=10)), ("vae", dict(x_dist="nbd"))])
def test_infer_ex_second_count_matrix_and_stripped_zero_inflation(model, kw):
  """llk of x_eval (the ORIGINAL counts) while the encoder reads x (the CORRUPTED counts), with and without the zero
  inflation; Monte-Carlo mean of the NB mean; importance weights -- all vs the float64 oracle on the same Philox noise."""
  from sisua_b200.engine import Engine
  from sisua_b200.posterior import apply_artificial_corruption
  G, B, S, seed = 300, 150, 4, 99
  cfg, flat, mov, batch = _setup(model, kw, G, B)
  x_org = batch["x"]
  x_cor = apply_artificial_corruption(x_org, dropout=0.3, retain_rate=0.2, seed=3)
  assert (x_cor != x_org).any() and (x_cor <= x_org).all()
  side = {k: v for k, v in batch.items() if k in ("y", "mask", "library")}
  eng = Engine(cfg, 0, flat_params=flat, bn_moving=mov)
  eps = _noise(cfg, B, S, seed, 7)
  ref = O.forward(cfg, Hh.oracle_params(cfg, flat), Hh.oracle_moving(cfg, mov), training=False, x=x_cor, **side, **eps)
  mu, th, pi = ref["mu"], ref["theta"], ref["pi_logit"]
  xo = torch.tensor(x_org, dtype=torch.float64)
  zi = cfg.x_dist == C.XDIST_ZINBD
  rec = (O.log_zinb_disp(xo, mu, th, pi) if zi else O.log_nb_disp(xo, mu, th)).sum(-1)
  imp = O.log_nb_disp(xo, mu, th).sum(-1)
  for strip, want in ((False, rec), (True, imp)):
    eng.set_infer_seed(seed, 7)
    out = eng.infer_ex(x_cor, x_eval=x_org, S=S, strip_zi=strip, want_mean_avg=True, want_logw=True, **side)
    _close(out["terms"][1].cpu().numpy().reshape(S, B), want.numpy(), what=f"llk of x_eval (strip_zi={strip})")
    _close(out["mean_avg"].cpu().numpy(), mu.mean(0).numpy(), atol=1e-7, what="Monte-Carlo mean of the NB mean")
  # importance weights log p(z) - log q(z | x) (+ library latent)
  z, zl, zs = ref["z"], ref["z_loc"], ref["z_scale"]
  logw = (torch.distributions.Normal(0., 1.).log_prob(z) - torch.distributions.Normal(zl, zs).log_prob(z)).sum(-1)
  if cfg.model_kind == C.MODEL_SCVI:
    lib = ref["l_loc"] + ref["l_scale"] * torch.tensor(eps["eps_l"], dtype=torch.float64)
    pm, pv = torch.tensor(side["library"][:, 0], dtype=torch.float64), torch.tensor(side["library"][:, 1], dtype=torch.float64)
    logw = logw + torch.distributions.Normal(pm, pv.sqrt()).log_prob(lib) - torch.distributions.Normal(ref["l_loc"], ref["l_scale"]).log_prob(lib)
  _close(out["logw"].cpu().numpy().reshape(S, B), logw.numpy(), rtol=2e-4, atol=2e-4, what="importance weights")
  eng.close()


@pytest.mark.parametrize("model,kw", [("vae", {}), ("scvi", {}), ("sisua", dict(n_proteins=10)), ("dca", {})])
def test_marginal_llk_matches_oracle(model, kw):
  from sisua_b200.engine import Engine
  G, B, S, seed = 200, 64, 50, 1234
  cfg, flat, mov, batch = _setup(model, kw, G, B, max_batch=4096) if False else _setup(model, kw, G, B)
  cfg = cfg.clone(max_batch=4096)
  eng = Engine(cfg, 0, flat_params=flat, bn_moving=mov)
  side = {k: v for k, v in batch.items() if k in ("y", "mask", "library")}
  eng.set_infer_seed(seed, 3)
  mllk, llk_x, llk_y = eng.marginal_llk(batch["x"], S=S, **side)
  if cfg.model_kind == C.MODEL_DCA:
    eps = {}
    ref = O.forward(cfg, Hh.oracle_params(cfg, flat), Hh.oracle_moving(cfg, mov), training=False, x=batch["x"])
    _close(mllk.cpu().numpy(), ref["llk_x"].numpy(), what="deterministic latent: marginal = llk")
    eng.close()
    return
  eps = _noise(cfg, B, S, seed, 3)
  ref = O.forward(cfg, Hh.oracle_params(cfg, flat), Hh.oracle_moving(cfg, mov), training=False, x=batch["x"], **side, **eps)
  z, zl, zs = ref["z"], ref["z_loc"], ref["z_scale"]
  logw = (torch.distributions.Normal(0., 1.).log_prob(z) - torch.distributions.Normal(zl, zs).log_prob(z)).sum(-1)
  if cfg.model_kind == C.MODEL_SCVI:
    lib = ref["l_loc"] + ref["l_scale"] * torch.tensor(eps["eps_l"], dtype=torch.float64)
    pm, pv = torch.tensor(side["library"][:, 0], dtype=torch.float64), torch.tensor(side["library"][:, 1], dtype=torch.float64)
    logw = logw + torch.distributions.Normal(pm, pv.sqrt()).log_prob(lib) - torch.distributions.Normal(ref["l_loc"], ref["l_scale"]).log_prob(lib)
  joint = ref["llk_x"] + logw
  if cfg.n_proteins > 0:
    joint = joint + cfg.alpha * torch.tensor(side["mask"], dtype=torch.float64) * ref["llk_y"]
  want = torch.logsumexp(joint, 0) - np.log(S)
  _close(mllk.cpu().numpy(), want.numpy(), rtol=1e-4, atol=1e-3, what="marginal llk")
  _close(llk_x.cpu().numpy(), (torch.logsumexp(ref["llk_x"], 0) - np.log(S)).numpy(), what="llk_x")
  if llk_y is not None:
    _close(llk_y.cpu().numpy(), (torch.logsumexp(ref["llk_y"], 0) - np.log(S)).numpy(), what="llk_y")
  eng.close()


@pytest.mark.parametrize("model,kw", [("vae", {}), ("scvi", {}), ("sisua", dict(n_proteins=10))])
def test_decode_only_equals_decoder_of_full_step(model, kw):
  """sisua_decode(z) reproduces the output parameters of the full inference step that produced z."""
  from sisua_b200.engine import Engine
  cfg, flat, mov, batch = _setup(model, kw, 260, 90)
  eng = Engine(cfg, 0, flat_params=flat, bn_moving=mov)
  out = eng.infer(want_mean=True, want_disp=True, want_pi=True, **batch)
  z = eng.debug_buffer("z", 90, cfg.n_latent)
  lib = None
  if model == "scvi":
    l_loc, l_scale = out["lib_loc"], out["lib_scale"]
    lib = l_loc + l_scale * torch.from_numpy(batch["eps_l"]).cuda()
  dec = eng.decode(z, lib)
  torch.cuda.synchronize()
  for k in ("mean", "disp", "pi_logit") + (("y_mean",) if model == "sisua" else ()):
    _close(dec[k].cpu().numpy(), out[k].cpu().numpy(), rtol=2e-5, atol=1e-7, what=k)
  eng.close()


@pytest.mark.parametrize("model,kw", [("vae", {}), ("scvi", {})])
def test_training_mode_forward_matches_oracle(model, kw):
  """`model(..., training=True)`: batch statistics + dropout masks from (seed, step), no gradients, parameters returned."""
  from sisua_b200.engine import Engine
  G, B = 200, 128
  cfg, flat, mov, batch = _setup(model, kw, G, B, input_dropout=0.3, enc_dropout=0.1)
  eng = Engine(cfg, 0, flat_params=flat, bn_moving=mov)
  out = eng.train_forward(seed=5, step=2, **batch)
  drop = Hh.oracle_dropout_masks(cfg, B, seed=5, step=2)
  ref = O.forward(cfg, Hh.oracle_params(cfg, flat), Hh.oracle_moving(cfg, mov), training=True, drop=drop, **batch)
  _close(out["terms"][0].cpu().numpy(), ref["elbo"].numpy(), what="elbo")
  _close(out["mean"].cpu().numpy(), ref["mu"].numpy(), atol=1e-7, what="mean")
  _close(out["pi_logit"].cpu().numpy(), ref["pi_logit"].numpy(), rtol=2e-4, atol=2e-5, what="pi")
  new_mov = PR.moving_to_dict(cfg, eng.bn_moving.cpu().numpy())
  for k, v in ref["new_moving"].items():
    _close(new_mov[k], v.numpy(), rtol=1e-4, atol=1e-6, what=k)
  assert float(eng.grads.abs().sum()) == 0.0          # no gradient was produced
  eng.close()


def test_dca_linear_latent_gradients():
  """RVmeta(.., 'linear'): the deterministic latent without the ReLU (what the reference coerces a stochastic latent to)."""
  from sisua_b200.engine import Engine
  cfg = C.make_step_config("dca", n_genes=150, max_batch=256, latent_linear=True)
  flat = Hh.randomize_norm_params(cfg, PR.init_flat_params(cfg)); mov = PR.init_bn_moving(cfg)
  batch = Hh.make_batch(cfg, 100)
  eng = Engine(cfg, 0, flat_params=flat, bn_moving=mov)
  terms, loss = eng.train_step(**batch)
  P = Hh.oracle_params(cfg, flat)
  for p in P.values():
    p.requires_grad_(True)
  ref = O.forward(cfg, P, Hh.oracle_moving(cfg, mov), training=True, **batch)
  ref["loss"].backward()
  assert (ref["z_loc"] < 0).any()                      # the linear latent really goes negative
  _close(terms[0].cpu().numpy(), ref["elbo"].detach().numpy(), what="elbo")
  got = eng.grads_dict()
  gmax = max(float(np.abs(p.grad.numpy()).max()) for p in P.values())
  for name, p in P.items():
    g = p.grad.numpy()       # (lat.b is exactly zero in exact arithmetic: BatchNorm follows the latent's decoder layer)
    assert np.abs(got[name] - g).max() <= 6e-3 * np.abs(g).max() + 1e-5 * gmax, name
  eng.close()


def test_streamed_predict_and_posterior_scores():
  """predict() keeps no [S, N, G] tensor: log_prob / imputed mean come from the fused kernels chunk by chunk and agree with
  the dense distributions; the Posterior returns the reference's keys."""
  from sisua_b200 import streamed as ST
  from sisua_b200.models import VAE, RVmeta, SingleCellData
  d = SY.realistic_counts(700, 120, 0, seed=11)
  sco = SingleCellData(d["x"], name="toy")
  train, test = sco.split(0.7)
  m = VAE(RVmeta(120, "zinbd", True, "transcriptomic"), max_batch=256, seed=4)       # 3 samples -> chunks of 85 cells
  m.fit(train, batch_size=64, epochs=3, learning_rate=2e-3)
  pX, qZ = m.predict(test, sample_shape=3, verbose=False)
  assert isinstance(pX, ST.StreamedIndependent) and isinstance(pX.distribution, D.ZeroInflated)
  assert isinstance(pX.distribution.count_distribution, D.NegativeBinomialDisp)
  N = len(test)
  assert tuple(pX.batch_shape) == (3, N) and tuple(pX.event_shape) == (120,) and tuple(qZ.batch_shape) == (N,)
  x = torch.from_numpy(test.X).cuda()
  lp = pX.log_prob(test.X)
  dense = pX.materialize()
  assert torch.allclose(lp, dense.log_prob(x), rtol=2e-4, atol=1e-2)
  assert torch.allclose(lp, pX.elbo_terms[1], rtol=1e-5, atol=1e-3)                   # same noise on every pass
  imp = ST.imputed_distribution(pX)
  assert torch.allclose(imp.log_prob(test.X), ST.imputed_distribution(dense).log_prob(x), rtol=2e-4, atol=1e-2)
  assert torch.allclose(pX.mean_over_samples(), dense.distribution.count_distribution.mean().mean(0), rtol=1e-5, atol=1e-6)
  post = m.create_posterior(test, sample_shape=3, batch_size=16)
  llk = post.cal_llk()
  assert set(llk) == {f"llk_transcriptomic_{a}_{b}" for a in ("imp", "rec") for b in ("org", "cor")}
  assert all(np.isfinite(v) for v in llk.values())
  sc = post.cal_imputation_scores()
  assert set(sc) == {"imputation_med", "imputation_mean", "imputation_std"} and all(np.isfinite(v) for v in sc.values())
  mm = post.cal_marginal_llk(sample_shape=20)
  assert set(mm) == {"transcriptomic_llk", "marginal_llk"} and mm["marginal_llk"] <= mm["transcriptomic_llk"] + 1e-3
  assert post.imputed.shape == (N, 120) and post.latents.shape == (N, 10)


def test_device_corruption_is_bit_exact_with_its_restatement():
  """sisua_corrupt_counts against oracle/philox.py:corrupt_counts on the same matrix: identical, for both distributions,
  with a strided source, in place, and with a count in the hundreds (many Philox calls for one entry)."""
  from oracle import philox as PH
  from sisua_b200 import config as C
  from sisua_b200.engine import Engine
  rng = np.random.default_rng(3)
  x = (rng.poisson(0.7, (777, 203)) * (rng.random((777, 203)) < 0.6)).astype(np.float32)
  x[3, 5] = 611.0; x[700, 202] = 65.0
  eng = Engine(C.make_step_config("vae", n_genes=203, n_latent=10, max_batch=64), 0, seed=1)
  for dist, seed in (("binomial", 7), ("uniform", (1 << 40) + 5)):
    got = eng.corrupt_counts(torch.from_numpy(x), 0.3, 0.2, dist, seed).cpu().numpy()
    np.testing.assert_array_equal(got, PH.corrupt_counts(x, 0.3, 0.2, dist, seed))
  wide = torch.zeros((777, 256), device="cuda"); wide[:, :203] = torch.from_numpy(x).cuda()
  view = wide[:, :203]
  eng.corrupt_counts(view, 0.3, 0.2, "binomial", 7, out=view)          # strided, in place
  np.testing.assert_array_equal(view.cpu().numpy(), PH.corrupt_counts(x, 0.3, 0.2, "binomial", 7))
  with pytest.raises(ValueError):
    eng.corrupt_counts(torch.from_numpy(x), 0.3, 0.2, "poisson", 7)
  eng.close()


def test_posterior_with_device_corruption():
  """Posterior(corrupt_on='device'): the test set is corrupted by sisua_corrupt_counts instead of the host routine."""
  from oracle import philox as PH
  from sisua_b200.models import VAE, RVmeta, SingleCellData
  from sisua_b200.posterior import Posterior
  d = SY.realistic_counts(400, 120, 0, seed=11)
  sco = SingleCellData(d["x"], name="toy")
  m = VAE(RVmeta(120, "zinbd", True, "transcriptomic"), max_batch=256, seed=4)
  m.fit(sco, batch_size=64, epochs=1, learning_rate=2e-3)
  post = Posterior(m, sco, dropout_rate=0.25, retain_rate=0.2, sample_shape=2, random_state=5, corrupt_on="device")
  np.testing.assert_array_equal(post.sco_corrupted.X, PH.corrupt_counts(sco.X, 0.25, 0.2, "binomial", 5))
  assert all(np.isfinite(v) for v in post.cal_imputation_scores().values())
  with pytest.raises(ValueError):
    Posterior(m, sco, corrupt_on="nowhere")
