"""CPU oracle of SISUA's minibatch ELBO train / infer step.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the arithmetic of this path lives in odin-ai==1.2.5 (reference
setup.py:27) on TensorFlow / TensorFlow-Probability.  None of the three is vendored
under /root/reference nor installable here (no network, Python 3.7 pin), and the
reference's tests hold no golden vectors for the path.  This file
restates the published algorithm and anchors it on the reference's own call sites:

  log1p input normalisation ............ sisua/models/single_cell_model.py:119-139
  encoder/decoder NetConf defaults ...... sisua/models/single_cell_model.py:74-86,
                                          configs/base.yaml:10-17
  latent RVmeta(10,'diag') + analytic KL  sisua/models/single_cell_model.py:77,91
  scVI graph (softmax scale, exp library,
    clip, parameter order) .............. sisua/models/scvi.py:88-171
  DCA deterministic latent ............... sisua/models/dca.py:16-28
  SISUA multitask label head ............. sisua/models/vae.py:19-44, configs/base.yaml:6,38-40
  minibatch dict (inputs, library, mask)  sisua/data/_single_cell_base.py:566-602
  optimiser defaults ..................... configs/base.yaml:45-50
  ZINB / NB log-likelihood (eps = 1e-8) .. scVI ``log_zinb_positive`` / ``log_nb_positive``
                                          (the formulas odin-ai's NegativeBinomialDisp /
                                          ZeroInflated follow)
  Keras Adam + per-variable clipnorm ..... SURVEY.md Appendix A

It is pinned by closed-form known-answer tests instead (tests/test_oracle_kat.py:
scipy.stats.nbinom, pmf normalisation, Gaussian KL, finite differences, a hand
computed Adam step).  Only tests/, __graft_entry__.smoke() and bench.py's CPU baseline
may import this module; the product (sisua_b200/) never does.

Everything is written with torch so the same code gives (a) the float64 truth used
for parity, (b) gradients via autograd, (c) the timed float32 CPU baseline.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

# enums duplicated on purpose (the oracle must not import the product package)
MODEL_VAE, MODEL_SCVI, MODEL_DCA, MODEL_SISUA = 0, 1, 2, 3
XDIST_ZINBD, XDIST_NBD, XDIST_ZINB, XDIST_NB = 0, 1, 2, 3
YDIST_NB, YDIST_NBD = 0, 1
ACT_SOFTPLUS, ACT_SOFTPLUS1, ACT_SOFTPLUS_P1, ACT_EXP, ACT_IDENTITY = 0, 1, 2, 3, 4
SOFTPLUS1_SHIFT = 0.5413248546129181
EPS = 1e-8


def activation(kind: int, x: torch.Tensor) -> torch.Tensor:
  """Q1 of SURVEY.md section 8a: candidate meanings of odin-ai's 'softplus1'."""
  if kind == ACT_SOFTPLUS:
    return F.softplus(x)
  if kind == ACT_SOFTPLUS1:
    return F.softplus(x + SOFTPLUS1_SHIFT)
  if kind == ACT_SOFTPLUS_P1:
    return F.softplus(x) + 1.0
  if kind == ACT_EXP:
    return torch.exp(x)
  if kind == ACT_IDENTITY:
    return x
  raise ValueError(kind)


# ----------------------------------------------------------------------------
# likelihoods
# ----------------------------------------------------------------------------
def log_nb_disp(x, mu, theta):
  """NB with mean / inverse-dispersion parameters, per element."""
  log_theta_mu = torch.log(theta + mu + EPS)
  return (theta * (torch.log(theta + EPS) - log_theta_mu) + x * (torch.log(mu + EPS) - log_theta_mu) +
          torch.lgamma(x + theta) - torch.lgamma(theta) - torch.lgamma(x + 1.0))


def log_zinb_disp(x, mu, theta, pi_logit):
  """Zero-inflated NB; ``pi_logit`` is the logit of the dropout probability."""
  sp_neg_pi = F.softplus(-pi_logit)
  log_theta_mu = torch.log(theta + mu + EPS)
  pi_theta_log = -pi_logit + theta * (torch.log(theta + EPS) - log_theta_mu)
  case_zero = F.softplus(pi_theta_log) - sp_neg_pi
  case_nonzero = (-sp_neg_pi + pi_theta_log + x * (torch.log(mu + EPS) - log_theta_mu) +
                  torch.lgamma(x + theta) - torch.lgamma(theta) - torch.lgamma(x + 1.0))
  return torch.where(x < EPS, case_zero, case_nonzero)


def log_nb_tfp(y, log_total_count, logits):
  """TFP NegativeBinomial(total_count = exp(a), logits = b) log-prob for real-valued y."""
  r = torch.exp(log_total_count)
  return (torch.lgamma(r + y) - torch.lgamma(r) - torch.lgamma(y + 1.0) + r * F.logsigmoid(-logits) +
          y * F.logsigmoid(logits))


def kl_diag_normal_std(loc, scale):
  return 0.5 * torch.sum(scale * scale + loc * loc - 1.0 - 2.0 * torch.log(scale), dim=-1)


def kl_normal_normal(loc, scale, prior_mean, prior_var):
  return (torch.log(torch.sqrt(prior_var) / scale) + (scale * scale + (loc - prior_mean) ** 2) /
          (2.0 * prior_var) - 0.5)


# ----------------------------------------------------------------------------
# network pieces
# ----------------------------------------------------------------------------
def _hidden_stack(cfg, P: Dict[str, torch.Tensor], prefix: str, n_layers: int, h, training: bool,
                  bn_moving: Optional[Dict[str, torch.Tensor]], new_moving: Dict[str, torch.Tensor],
                  drop: Optional[Dict[str, torch.Tensor]], rate: float, trace: Optional[Dict[str, torch.Tensor]] = None):
  """Dense(no bias when BN) -> BatchNorm -> ReLU -> Dropout, n_layers times."""
  for i in range(n_layers):
    name = f"{prefix}.{i}"
    a = h @ P[name + ".W"].T
    if cfg.batchnorm:
      if training:
        mean = a.mean(dim=0)
        var = a.var(dim=0, unbiased=False)
        if bn_moving is not None:
          m = cfg.bn_momentum
          new_moving[name + ".mean"] = m * bn_moving[name + ".mean"] + (1.0 - m) * mean.detach()
          new_moving[name + ".var"] = m * bn_moving[name + ".var"] + (1.0 - m) * var.detach()
      else:
        mean, var = bn_moving[name + ".mean"], bn_moving[name + ".var"]
      a = (a - mean) / torch.sqrt(var + cfg.bn_eps) * P[name + ".gamma"] + P[name + ".beta"]
    else:
      a = a + P[name + ".b"]
    if trace is not None:       # ReLU inputs, for tests that need to know how close to zero they come
      trace[name] = a.detach()
    h = torch.relu(a)
    if training and rate > 0.0:
      if drop is None or name not in drop:
        raise ValueError(f"training with dropout needs an explicit mask for layer {name}")
      h = h * drop[name] / (1.0 - rate)
  return h


def forward(cfg, P: Dict[str, torch.Tensor], bn_moving: Optional[Dict[str, torch.Tensor]], x, y=None,
            library=None, mask=None, eps_z=None, eps_l=None, training: bool = False,
            drop: Optional[Dict[str, torch.Tensor]] = None, trace: Optional[Dict[str, torch.Tensor]] = None):
  """One ELBO evaluation.  x [B,G]; y [B,P]; library [B,2] = (mean, var) of log-library;
  mask [B] in {0,1}; eps_z [S,B,Z] or [B,Z]; eps_l [S,B] or [B].
  Returns a dict of per-cell terms (leading sample axis S kept when eps has one)."""
  dt = P["out.W"].dtype
  x = torch.as_tensor(x, dtype=dt)
  B, G = x.shape
  H, Z = cfg.n_hidden, cfg.n_latent
  new_moving: Dict[str, torch.Tensor] = {}
  xt = torch.log1p(x) if cfg.log_norm else x
  if training and cfg.input_dropout > 0.0:
    xt = xt * drop["input"] / (1.0 - cfg.input_dropout)
  h = _hidden_stack(cfg, P, "enc", cfg.n_enc_layers, xt, training, bn_moving, new_moving, drop,
                    cfg.enc_dropout, trace)
  p = h @ P["lat.W"].T + P["lat.b"]
  out = {}
  deterministic = cfg.model_kind == MODEL_DCA
  if deterministic:
    z_loc = p if getattr(cfg, "latent_linear", 0) else torch.relu(p)      # RVmeta(.., 'relu') default, 'linear' when coerced (dca.py:16-27)
    z_scale = torch.zeros_like(z_loc)
    z = z_loc
    kl_z = torch.zeros(B, dtype=dt)
    sample_axis = False
  else:
    z_loc = p[:, :Z]
    z_scale = activation(cfg.scale_act, p[:, Z:])
    e = torch.as_tensor(eps_z, dtype=dt)
    sample_axis = e.dim() == 3
    z = z_loc + z_scale * e            # broadcasts over S
    kl_z = kl_diag_normal_std(z_loc, z_scale)
  kl_l = torch.zeros(B, dtype=dt)
  lib = None
  if cfg.model_kind == MODEL_SCVI:
    hl = _hidden_stack(cfg, P, "encl", cfg.n_encl_layers, xt, training, bn_moving, new_moving, drop,
                       cfg.encl_dropout, trace)
    pl = hl @ P["lib.W"].T + P["lib.b"]
    l_loc = pl[:, 0]
    l_scale = activation(cfg.scale_act, pl[:, 1])
    el = torch.as_tensor(eps_l, dtype=dt)
    lib = l_loc + l_scale * el         # [B] or [S,B]
    library = torch.as_tensor(library, dtype=dt)
    kl_l = kl_normal_normal(l_loc, l_scale, library[:, 0], library[:, 1])
    out["l_loc"], out["l_scale"] = l_loc, l_scale
  S = z.shape[0] if sample_axis else 1
  zf = z.reshape(S * B, Z)
  # training-mode BN in the decoder sees all S*B rows (the reference flattens the MC
  # axis before the decoder: scvi.py:118-127)
  d = _hidden_stack(cfg, P, "dec", cfg.n_dec_layers, zf, training, bn_moving, new_moving, drop,
                    cfg.dec_dropout, trace)
  if trace is not None and trace.get("__hidden_only__"):      # tests: ReLU inputs only, skip the gene-sized likelihood
    return {"d": d}
  o = d @ P["out.W"].T + P["out.b"]
  xa = x.repeat(S, 1) if S > 1 else x
  a, b = o[:, :G], o[:, G:2 * G]
  if cfg.model_kind == MODEL_SCVI:
    scale = torch.clamp(torch.softmax(a, dim=1), 1e-7, 1.0 - 1e-7)
    Lc = torch.clamp(lib.reshape(S * B, 1), 0.0, cfg.clip_library)
    mu = torch.exp(Lc) * scale
    theta = torch.exp(b)
    if cfg.scvi_reapply_act:           # Q2, literal reading of projection=False
      mu = activation(cfg.mean_act, mu)
      theta = activation(cfg.disp_act, theta)
  elif cfg.x_dist in (XDIST_ZINB, XDIST_NB):
    # TFP NegativeBinomial(total_count = exp(a), logits = b) (tests/test_singlecell_models.py:60-80): reported through its
    # (mean, inverse dispersion) = (exp(a + b), exp(a))
    theta = torch.exp(a)
    mu = torch.exp(a + b)
  else:
    mu = activation(cfg.mean_act, a)
    theta = activation(cfg.disp_act, b)
  if cfg.x_dist in (XDIST_ZINB, XDIST_NB):
    base = log_nb_tfp(xa, a, b)
    if cfg.x_dist == XDIST_ZINB:
      pi = o[:, 2 * G:3 * G]
      llk_x = torch.where(xa < EPS, F.softplus(base - pi) - F.softplus(-pi), base - F.softplus(pi)).sum(dim=1)
    else:
      pi = None
      llk_x = base.sum(dim=1)
  elif cfg.x_dist == XDIST_ZINBD:
    pi = o[:, 2 * G:3 * G]
    llk_x = log_zinb_disp(xa, mu, theta, pi).sum(dim=1)
  else:
    pi = None
    llk_x = log_nb_disp(xa, mu, theta).sum(dim=1)
  llk_y = torch.zeros(S * B, dtype=dt)
  if cfg.n_proteins > 0:
    Pn = cfg.n_proteins
    y = torch.as_tensor(y, dtype=dt)
    ya = y.repeat(S, 1) if S > 1 else y
    py = d @ P["y.W"].T + P["y.b"]
    if cfg.y_dist == YDIST_NB:
      llk_y = log_nb_tfp(ya, py[:, :Pn], py[:, Pn:]).sum(dim=1)
      y_mean = torch.exp(py[:, :Pn]) * torch.exp(py[:, Pn:])
    else:
      ymu = activation(cfg.mean_act, py[:, :Pn])
      yth = activation(cfg.disp_act, py[:, Pn:])
      llk_y = log_nb_disp(ya, ymu, yth).sum(dim=1)
      y_mean = ymu
    out["y_mean"] = y_mean.reshape(S, B, Pn) if sample_axis else y_mean
  m = torch.zeros(B, dtype=dt) if mask is None else torch.as_tensor(mask, dtype=dt).reshape(B)
  llk_x = llk_x.reshape(S, B)
  llk_y = llk_y.reshape(S, B)
  if cfg.mask_norm == 1 and cfg.n_proteins > 0:
    # Q3 alternative: average the supervised term over labelled cells only
    w_y = m * (B / torch.clamp(m.sum(), min=1.0))
  else:
    w_y = m
  elbo = llk_x + cfg.alpha * w_y * llk_y - cfg.beta * (kl_z + kl_l)
  loss = -elbo.mean()

  def shp(t, trailing):
    t = t.reshape((S, B) + trailing)
    return t if sample_axis else t[0]

  out.update(dict(
      loss=loss, elbo=elbo if sample_axis else elbo[0], llk_x=llk_x if sample_axis else llk_x[0],
      llk_y=llk_y if sample_axis else llk_y[0], kl_z=kl_z, kl_l=kl_l, z_loc=z_loc, z_scale=z_scale,
      z=z, d=shp(d, (H,)), mu=shp(mu, (G,)), theta=shp(theta, (G,)),
      pi_logit=None if pi is None else shp(pi, (G,)), new_moving=new_moving))
  return out


# ----------------------------------------------------------------------------
# optimiser: Keras/TF Adam + clipnorm
# ----------------------------------------------------------------------------
def adam_update(P, grads, m, v, t: int, lr=1e-3, beta1=0.9, beta2=0.999, eps_hat=1e-7, clipnorm=100.0,
                clip_mode=0):
  """In-place TF-style Adam on dicts of tensors. t is the 1-based step index."""
  with torch.no_grad():
    if clipnorm and clipnorm > 0 and clip_mode == 1:
      gn = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values()))
      gscale = float(min(1.0, clipnorm / max(float(gn), 1e-30)))
    lr_t = lr * math.sqrt(1.0 - beta2 ** t) / (1.0 - beta1 ** t)
    for k, g in grads.items():
      if clipnorm and clipnorm > 0:
        if clip_mode == 0:
          n = float(torch.sqrt((g.double() ** 2).sum()))
          g = g * min(1.0, clipnorm / max(n, 1e-30))
        else:
          g = g * gscale
      m[k].mul_(beta1).add_(g, alpha=1.0 - beta1)
      v[k].mul_(beta2).addcmul_(g, g, value=1.0 - beta2)
      P[k].sub_(lr_t * m[k] / (torch.sqrt(v[k]) + eps_hat))


def train_step(cfg, P, bn_moving, m, v, t, batch, lr=1e-3, clipnorm=100.0, drop=None):
  """forward + autograd backward + Adam; mutates P, m, v, bn_moving. Returns (out, grads)."""
  for p in P.values():
    p.requires_grad_(True)
    p.grad = None
  out = forward(cfg, P, bn_moving, training=True, drop=drop, **batch)
  out["loss"].backward()
  grads = {k: (p.grad.detach().clone() if p.grad is not None else torch.zeros_like(p))
           for k, p in P.items()}
  for p in P.values():
    p.requires_grad_(False)
    p.grad = None
  adam_update(P, grads, m, v, t, lr=lr, clipnorm=clipnorm, clip_mode=cfg.clip_mode)
  if bn_moving is not None:
    for k, val in out["new_moving"].items():
      bn_moving[k] = val
  return out, grads


# ----------------------------------------------------------------------------
# derived quantities downstream code consumes (sisua/analysis/posterior.py:210-220)
# ----------------------------------------------------------------------------
def imputed_mean(out):
  """NB mean without zero inflation, averaged over MC samples when present."""
  mu = out["mu"]
  return mu.mean(dim=0) if mu.dim() == 3 else mu


def reconstructed_mean(out):
  mu, pi = out["mu"], out["pi_logit"]
  r = mu if pi is None else torch.sigmoid(-pi) * mu
  return r.mean(dim=0) if r.dim() == 3 else r


def library_size_stats(X: np.ndarray):
  """sisua/data/utils.py:231-263 — dataset-level mean / variance of log library size,
  broadcast to every cell -> [N, 2]."""
  total = np.asarray(X, dtype=np.float64).sum(axis=1)
  lc = np.log(total + 1e-8)
  return np.stack([np.full(X.shape[0], lc.mean()), np.full(X.shape[0], lc.var())], axis=1).astype(np.float32)
