"""Shared test plumbing: product param buffers -> oracle dicts, seeded batches."""
import numpy as np
import torch

from sisua_b200 import config as C
from sisua_b200 import params as PR
from sisua_b200 import synthetic as SY


def oracle_params(cfg, flat, dtype=torch.float64):
  return {k: torch.tensor(np.array(v), dtype=dtype) for k, v in PR.flat_to_dict(cfg, flat).items()}


def oracle_moving(cfg, moving, dtype=torch.float64):
  return {k: torch.tensor(np.array(v), dtype=dtype) for k, v in PR.moving_to_dict(cfg, moving).items()}


def randomize_norm_params(cfg, flat, seed=3):
  """Make gamma/beta/biases non-trivial so parity tests exercise them."""
  rng = np.random.default_rng(seed)
  d = PR.flat_to_dict(cfg, flat)
  for e in C.param_layout(cfg)[0]:
    if e.kind == "gamma":
      d[e.name][...] = rng.uniform(0.5, 1.5, size=e.shape).astype(np.float32)
    elif e.kind in ("beta", "bias"):
      d[e.name][...] = rng.normal(0, 0.1, size=e.shape).astype(np.float32)
  return flat


def make_batch(cfg, B, seed=0, S=None, preset_stats=(6.42, 0.28), stress=False):
  G, P, Z = cfg.n_genes, cfg.n_proteins, cfg.n_latent
  if stress:
    data = SY.stress_counts(B, G, P, seed=SY.DATA_SEED + seed)
  else:
    data = SY.realistic_counts(B, G, P, preset_stats[0], preset_stats[1], seed=SY.DATA_SEED + seed)
  rng = np.random.default_rng(1000 + seed)
  shape = (B,) if S is None else (S, B)
  batch = dict(x=data["x"])
  if P > 0:
    batch["y"] = data["y"]
    m = (rng.random(B) < 0.3).astype(np.uint8)
    m[0] = 1
    batch["mask"] = m
  if cfg.model_kind == C.MODEL_SCVI:
    batch["library"] = SY.library_stats(data["x"])
    batch["eps_l"] = rng.standard_normal(shape).astype(np.float32)
  if cfg.model_kind != C.MODEL_DCA:
    batch["eps_z"] = rng.standard_normal(shape + (Z,)).astype(np.float32)
  return batch


def oracle_dropout_masks(cfg, B, seed, step, dtype=torch.float64):
  """The masks the CUDA step regenerates from (seed, step): stream 0 = input, 1+u = hidden unit u
  (units numbered enc, encl, dec — sisua_b200/csrc/abi.cu stat_index)."""
  from oracle.philox import dropout_mask
  units = [f"enc.{i}" for i in range(cfg.n_enc_layers)]
  rates = [cfg.enc_dropout] * cfg.n_enc_layers
  if cfg.model_kind == C.MODEL_SCVI:
    units += [f"encl.{i}" for i in range(cfg.n_encl_layers)]
    rates += [cfg.encl_dropout] * cfg.n_encl_layers
  units += [f"dec.{i}" for i in range(cfg.n_dec_layers)]
  rates += [cfg.dec_dropout] * cfg.n_dec_layers
  drop = {}
  if cfg.input_dropout > 0:
    drop["input"] = torch.tensor(dropout_mask(B, cfg.n_genes, cfg.input_dropout, seed, step, 0), dtype=dtype)
  for u, (name, rate) in enumerate(zip(units, rates)):
    if rate > 0:
      drop[name] = torch.tensor(dropout_mask(B, cfg.n_hidden, rate, seed, step, 1 + u), dtype=dtype)
  return drop


def separate_relu_ties(cfg, flat, mov, batch, drop=None, training=True):
  """Nudges the beta / bias of every hidden unit (by less than the spacing of its inputs, ~1e-4) so that no ReLU input of
  THIS batch lies close to zero, layer by layer in forward order; returns the smallest |ReLU input| margin reached.

  Why: with millions of ReLU inputs per step a few always sit within float32 rounding of zero.  Such a unit is "on" in one
  arithmetic and "off" in another, and the gradients then differ by that cell's whole contribution (1e-3 .. 1e-2 of a
  tensor's largest entry) -- the float32 and float64 ORACLES disagree with each other in exactly that way.  It is a
  property of ReLU in finite precision, not of the kernels, and it would make a full-batch gradient comparison random."""
  from oracle import step_oracle as O
  names = [f"enc.{i}" for i in range(cfg.n_enc_layers)]
  if cfg.model_kind == C.MODEL_SCVI:
    names += [f"encl.{i}" for i in range(cfg.n_encl_layers)]
  names += [f"dec.{i}" for i in range(cfg.n_dec_layers)]
  d = PR.flat_to_dict(cfg, flat)         # views into `flat`
  key = ".beta" if cfg.batchnorm else ".b"
  margin = np.inf
  for name in names:
    trace = {"__hidden_only__": True}
    with torch.no_grad():
      O.forward(cfg, oracle_params(cfg, flat), oracle_moving(cfg, mov), training=training, drop=drop, trace=trace, **batch)
    a = trace[name].numpy()              # [rows, H] ReLU inputs of this unit under the current parameters
    for c in range(a.shape[1]):
      v = np.sort(a[:, c])
      # shifting beta by -m moves every input by -m: m = the centre of the widest gap among the inputs closest to zero
      k = np.searchsorted(v, 0.0)
      w = v[max(0, k - 8):min(v.size, k + 8)]
      if w.size < 2:
        continue
      gaps = np.diff(w)
      j = int(np.argmax(gaps))
      d[name + key][c] -= np.float32(0.5 * (w[j] + w[j + 1]))
      margin = min(margin, 0.5 * gaps[j])
  return float(margin)
