"""Keras-side view of the weights: the flat fp32 parameter buffer re-expressed as the variable
list a Keras model of the same architecture holds, and back.

The reference checkpoints with ``keras.Model.save_weights`` in TF format (sisua/models/single_cell_model.py:295-306) and
reloads with ``load_weights`` (:283-293, sisua/models/__init__.py:30-38).  A TF checkpoint can only be read with
TensorFlow, which this image does not have, so the exchange format here is the one any TF install produces in one line::

    np.savez(path, **{v.name: v.numpy() for v in model.variables + model.optimizer.variables()})

i.e. an ``.npz`` keyed by Keras variable names.  This module maps between those names / Keras tensor conventions and
the flat buffer:

* ``Dense.kernel`` is ``[in_features, out_features]`` -- the transpose of the ``[out, in]`` (K-major) layout of
  ``config.param_layout``;
* a layer with BatchNormalization has no Dense bias (``NetConf(batchnorm=True)`` builds ``Dense(use_bias=False)`` +
  ``BatchNormalization``): ``gamma, beta, moving_mean, moving_variance``;
* the output heads are ONE ``Dense(3 G)`` whose columns are ``mean | dispersion | dropout logit`` (the order
  ``DistributionDense`` splits its parameters in), the latent head one ``Dense(2 Z)`` (``loc | scale``);
* Adam slots follow Keras: ``Adam/<variable>/m``, ``Adam/<variable>/v``, ``Adam/iter``.

The scope names (``encoder/dense_0/...``) follow Keras's default numbering inside the reference's sub-networks; odin-ai's
exact scopes cannot be checked here (the package is absent), so every function takes ``name_map`` -- ``{our default
name: name in the file}`` -- and ``match_by_shape`` pairs a foreign variable list with ours by order and shape."""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np

from .config import StepConfig, bn_layer_names, param_layout
from .params import dict_to_flat, flat_to_dict

_SCOPES = {"enc": "encoder", "encl": "encoder_l", "dec": "decoder", "lat": "latents", "lib": "library", "y": "labels", "out": "outputs"}


def keras_name(our: str) -> str:
  """``enc.1.W -> encoder/dense_1/kernel:0``, ``dec.0.gamma -> decoder/batch_normalization/gamma:0``, ..."""
  parts = our.split(".")
  scope = _SCOPES[parts[0]]
  if len(parts) == 3:
    idx, leaf = int(parts[1]), parts[2]
    sfx = "" if idx == 0 else f"_{idx}"
    if leaf == "W":
      return f"{scope}/dense{sfx}/kernel:0"
    if leaf == "b":
      return f"{scope}/dense{sfx}/bias:0"
    leaf = {"gamma": "gamma", "beta": "beta", "mean": "moving_mean", "var": "moving_variance"}[leaf]
    return f"{scope}/batch_normalization{sfx}/{leaf}:0"
  leaf = {"W": "kernel", "b": "bias"}[parts[1]]
  return f"{scope}/dense/{leaf}:0"


def variable_names(cfg: StepConfig) -> List[Tuple[str, str, Tuple[int, ...]]]:
  """[(our name, Keras name, Keras shape)] of every model variable, trainable ones first (layout order), then the
  BatchNorm moving statistics."""
  out = []
  for e in param_layout(cfg)[0]:
    shape = (e.shape[1], e.shape[0]) if len(e.shape) == 2 else tuple(e.shape)
    out.append((e.name, keras_name(e.name), shape))
  for l in bn_layer_names(cfg):
    for leaf in ("mean", "var"):
      out.append((f"{l}.{leaf}", keras_name(f"{l}.{leaf}"), (cfg.n_hidden,)))
  return out


def to_keras(cfg: StepConfig, flat: np.ndarray, bn_moving: np.ndarray, adam_m: Optional[np.ndarray] = None,
             adam_v: Optional[np.ndarray] = None, step: Optional[int] = None, name_map: Optional[Dict[str, str]] = None) -> Dict[str, np.ndarray]:
  """Keras-named, Keras-shaped arrays of the model (and, if given, of the Adam slots)."""
  nm = name_map or {}
  rename = lambda our: nm.get(our, keras_name(our))
  out: Dict[str, np.ndarray] = {}

  for name, t in flat_to_dict(cfg, np.asarray(flat, dtype=np.float32)).items():
    out[rename(name)] = np.ascontiguousarray(t.T if t.ndim == 2 else t).astype(np.float32)
  for i, l in enumerate(bn_layer_names(cfg)):
    out[rename(f"{l}.mean")] = np.array(bn_moving[i, 0], dtype=np.float32)
    out[rename(f"{l}.var")] = np.array(bn_moving[i, 1], dtype=np.float32)
  for slot, buf in (("m", adam_m), ("v", adam_v)):
    if buf is None:
      continue
    for name, t in flat_to_dict(cfg, np.asarray(buf, dtype=np.float32)).items():
      out[f"Adam/{rename(name)[:-2]}/{slot}:0"] = np.ascontiguousarray(t.T if t.ndim == 2 else t).astype(np.float32)
  if step is not None:
    out["Adam/iter:0"] = np.int64(step)
  return out


def from_keras(cfg: StepConfig, arrays: Dict[str, np.ndarray], name_map: Optional[Dict[str, str]] = None, strict: bool = True):
  """(flat, bn_moving, adam_m | None, adam_v | None, step | None) from Keras-named arrays.  Shapes are checked against
  the architecture; with ``strict`` a missing model variable raises ``KeyError`` (Adam slots are optional)."""
  nm = name_map or {}
  rename = lambda our: nm.get(our, keras_name(our))
  entries = param_layout(cfg)[0]

  def take(key, our, shape, what):
    if key not in arrays:
      if strict:
        raise KeyError(f"{what} '{key}' (for {our}) is not in the file; present: {sorted(arrays)[:8]}...")
      return None
    a = np.asarray(arrays[key], dtype=np.float32)
    if tuple(a.shape) != tuple(shape):
      raise ValueError(f"'{key}' has shape {tuple(a.shape)}, the architecture needs {tuple(shape)} for {our}")
    return a

  def gather(prefix_fn, what, need):
    tensors = {}
    for e in entries:
      kshape = (e.shape[1], e.shape[0]) if len(e.shape) == 2 else tuple(e.shape)
      a = take(prefix_fn(e.name), e.name, kshape, what) if need or prefix_fn(e.name) in arrays else None
      if a is None:
        return None
      tensors[e.name] = a.T if a.ndim == 2 else a
    return dict_to_flat(cfg, tensors)

  flat = gather(rename, "variable", True)
  names = bn_layer_names(cfg)
  moving = np.zeros((max(len(names), 1), 2, cfg.n_hidden), dtype=np.float32)
  moving[:, 1] = 1.0
  for i, l in enumerate(names):
    for j, leaf in enumerate(("mean", "var")):
      a = take(rename(f"{l}.{leaf}"), f"{l}.{leaf}", (cfg.n_hidden,), "moving statistic")
      if a is not None:
        moving[i, j] = a
  m = gather(lambda n: f"Adam/{rename(n)[:-2]}/m:0", "Adam slot", False)
  v = gather(lambda n: f"Adam/{rename(n)[:-2]}/v:0", "Adam slot", False)
  step = int(arrays["Adam/iter:0"]) if "Adam/iter:0" in arrays else None
  return flat, moving, m, v, step


def match_by_shape(cfg: StepConfig, foreign: List[Tuple[str, Tuple[int, ...]]]) -> Dict[str, str]:
  """``name_map`` for a foreign variable list ``[(name, shape)]`` (e.g. ``[(v.name, v.shape) for v in model.variables]``)
  whose scopes differ from the defaults: every variable of ours takes the first unused foreign variable of its Keras
  shape, in order -- Keras lists variables in construction order, which is the layout order here.  Raises if a variable
  finds no partner."""
  used, out = set(), {}
  for our, _, shape in variable_names(cfg):
    leaf = keras_name(our).rsplit("/", 1)[1]
    for i, (name, fshape) in enumerate(foreign):
      if i in used or tuple(fshape) != tuple(shape) or name.rsplit("/", 1)[-1] != leaf:
        continue
      used.add(i); out[our] = name
      break
    else:
      raise KeyError(f"no foreign variable of shape {shape} named */{leaf} left for {our}")
  return out
