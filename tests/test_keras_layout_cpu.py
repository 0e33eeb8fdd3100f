"""Keras-named weight exchange (sisua_b200/keras_layout.py): host logic, no GPU."""
import numpy as np
import pytest

from sisua_b200 import config as C
from sisua_b200 import keras_layout as K
from sisua_b200 import params as PR


def _cfg(model):
  kw = dict(n_proteins=7) if model == "sisua" else {}
  return C.make_step_config(model, n_genes=50, n_latent=10, max_batch=64, **kw)


@pytest.mark.parametrize("model", ["vae", "scvi", "sisua", "dca"])
def test_round_trip_and_keras_conventions(model):
  cfg = _cfg(model)
  rng = np.random.default_rng(0)
  total = C.param_layout(cfg)[1]
  flat = PR.dict_to_flat(cfg, {e.name: rng.normal(size=e.shape) for e in C.param_layout(cfg)[0]})
  m = PR.dict_to_flat(cfg, {e.name: rng.normal(size=e.shape) for e in C.param_layout(cfg)[0]})
  v = np.abs(PR.dict_to_flat(cfg, {e.name: rng.normal(size=e.shape) for e in C.param_layout(cfg)[0]}))
  mov = rng.normal(size=PR.init_bn_moving(cfg).shape).astype(np.float32)
  arrays = K.to_keras(cfg, flat, mov, m, v, step=17)
  # Dense kernels are [in, out]; the three output heads are the column blocks of one Dense(3 G)
  assert arrays["encoder/dense/kernel:0"].shape == (50, 64)
  assert arrays["outputs/dense/kernel:0"].shape == (64, cfg.n_out_heads * 50)
  assert arrays["latents/dense/kernel:0"].shape[0] == 64
  w = PR.flat_to_dict(cfg, flat)
  np.testing.assert_array_equal(arrays["outputs/dense/kernel:0"][:, 50:100], w["out.W"][50:100].T)
  assert "encoder/batch_normalization_1/moving_variance:0" in arrays and "Adam/encoder/dense/kernel/m:0" in arrays
  f2, mov2, m2, v2, step = K.from_keras(cfg, arrays)
  assert f2.shape == (total,) and step == 17
  np.testing.assert_array_equal(f2, flat); np.testing.assert_array_equal(m2, m); np.testing.assert_array_equal(v2, v)
  np.testing.assert_array_equal(mov2[:len(C.bn_layer_names(cfg))], mov[:len(C.bn_layer_names(cfg))])
  # without optimiser slots
  only = {k: a for k, a in arrays.items() if not k.startswith("Adam/")}
  f3, _, m3, v3, step3 = K.from_keras(cfg, only)
  np.testing.assert_array_equal(f3, flat)
  assert m3 is None and v3 is None and step3 is None


def test_foreign_scopes_missing_and_misshapen_variables():
  cfg = _cfg("vae")
  flat = PR.init_flat_params(cfg); mov = PR.init_bn_moving(cfg)
  arrays = K.to_keras(cfg, flat, mov)
  # a checkpoint whose scopes differ: pair by order + shape + leaf name
  foreign = {k.replace("encoder/", "vae/enc_net/").replace("outputs/", "rna/"): a for k, a in arrays.items()}
  with pytest.raises(KeyError):
    K.from_keras(cfg, foreign)
  nm = K.match_by_shape(cfg, [(k, a.shape) for k, a in foreign.items()])
  assert nm["enc.0.W"] == "vae/enc_net/dense/kernel:0" and nm["out.b"] == "rna/dense/bias:0"
  f2, mov2, _, _, _ = K.from_keras(cfg, foreign, name_map=nm)
  np.testing.assert_array_equal(f2, flat)
  bad = dict(arrays); bad["decoder/dense/kernel:0"] = np.zeros((11, 64), np.float32)
  with pytest.raises(ValueError):
    K.from_keras(cfg, bad)
  fewer = dict(arrays); del fewer["latents/dense/bias:0"]
  with pytest.raises(KeyError):
    K.from_keras(cfg, fewer)
