"""Flat fp32 parameter buffer helpers (host side, numpy).

The reference gets its initial weights from Keras defaults inside odin-ai's NetConf /
DenseDistribution (glorot-uniform kernels, zero biases, BN gamma=1 beta=0, moving
mean=0 var=1); seeds follow sisua/train.py:24-25 (SEED = 8)."""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np

from .config import StepConfig, bn_layer_names, param_layout


def init_flat_params(cfg: StepConfig, seed: int = 8) -> np.ndarray:
  entries, total = param_layout(cfg)
  rng = np.random.default_rng(seed)
  flat = np.zeros(total, dtype=np.float32)
  for e in entries:
    if e.kind == "weight":
      rows, cols = e.shape
      limit = np.sqrt(6.0 / (e.fan_in + e.fan_out))
      w = rng.uniform(-limit, limit, size=(rows, cols)).astype(np.float32)
      view = flat[e.offset:e.offset + rows * e.ld].reshape(rows, e.ld)
      view[:, :cols] = w
    elif e.kind == "gamma":
      flat[e.offset:e.offset + e.shape[0]] = 1.0
  return flat


def init_bn_moving(cfg: StepConfig) -> np.ndarray:
  """[n_bn_layers, 2, H]: moving mean (0) and moving variance (1)."""
  n = len(bn_layer_names(cfg))
  buf = np.zeros((max(n, 1), 2, cfg.n_hidden), dtype=np.float32)
  buf[:, 1, :] = 1.0
  return buf


def flat_to_dict(cfg: StepConfig, flat: np.ndarray) -> Dict[str, np.ndarray]:
  """Views (no copy) of every tensor in its logical shape, padding columns stripped."""
  entries, _ = param_layout(cfg)
  out = {}
  for e in entries:
    if len(e.shape) == 2:
      rows, cols = e.shape
      out[e.name] = flat[e.offset:e.offset + rows * e.ld].reshape(rows, e.ld)[:, :cols]
    else:
      out[e.name] = flat[e.offset:e.offset + e.shape[0]]
  return out


def dict_to_flat(cfg: StepConfig, tensors: Dict[str, np.ndarray]) -> np.ndarray:
  entries, total = param_layout(cfg)
  flat = np.zeros(total, dtype=np.float32)
  views = flat_to_dict(cfg, flat)
  for e in entries:
    views[e.name][...] = np.asarray(tensors[e.name], dtype=np.float32)
  return flat


def moving_to_dict(cfg: StepConfig, moving: np.ndarray) -> Dict[str, np.ndarray]:
  out = {}
  for i, name in enumerate(bn_layer_names(cfg)):
    out[name + ".mean"] = moving[i, 0]
    out[name + ".var"] = moving[i, 1]
  return out


def n_trainable(cfg: StepConfig) -> int:
  entries, _ = param_layout(cfg)
  return int(sum(int(np.prod(e.shape)) for e in entries))
