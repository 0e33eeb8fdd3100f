"""Drop-in surface of sisua.models on the GPU: the assertions mirror what the reference's own tests check
(tests/test_singlecell_models.py:28-36,93-188 class registry, loss decreasing, predict output types/shapes;
tests/test_save_load_model.py:195-201 latent statistics identical after save -> load)."""
import os

import numpy as np
import pytest
import torch

from sisua_b200 import distributions as D
from sisua_b200 import synthetic as SY
from sisua_b200.models import (SCVI, SISUA, VAE, DeepCountAutoencoder, NetConf, RVmeta, SingleCellData,
                               SingleCellModel, get_all_models, get_model, load_model)


def _data(n=512, g=60, p=0, labels_percent=0.0):
  d = SY.realistic_counts(n, g, p, seed=123)
  return SingleCellData(d["x"], d.get("y"), name="toy", labels_percent=labels_percent)


def _make(name, g=60, p=6, **kw):
  rna = RVmeta(g, 'zinbd', True, 'transcriptomic')
  if name == "sisua":
    return SISUA(rna, RVmeta(p, 'nb', True, 'proteomic'), max_batch=4096, **kw)
  return get_model(name)(rna, max_batch=4096, **kw)


def test_registry_and_ids():
  ids = {m.id for m in get_all_models()}
  assert {"vae", "sisua", "scvi", "dca", "scm"} <= ids
  assert get_model("dca") is DeepCountAutoencoder and get_model("SCVI") is SCVI and get_model(VAE) is VAE
  with pytest.raises(RuntimeError):
    get_model("nope")


def test_fit_requires_metadata_and_posterior_requires_fit():
  m = VAE(RVmeta(60, 'zinbd', True, 'transcriptomic'))
  with pytest.raises(RuntimeError):
    m.fit(np.zeros((64, 60), dtype=np.float32))
  with pytest.raises(RuntimeError):
    m.create_posterior()
  with pytest.raises(AssertionError):
    m.predict(np.zeros((4, 60), dtype=np.float32), device="TPU")


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["vae", "dca", "scvi", "sisua"])
@pytest.mark.parametrize("gemm_mode", [0, 1])
def test_fit_predict_roundtrip(name, gemm_mode, tmp_path):
  sco = _data(p=6 if name == "sisua" else 0, labels_percent=0.3)
  train, test = sco.split(0.8)
  model = _make(name, gemm_mode=gemm_mode, seed=1)
  assert not model.is_fitted
  model.fit(train, valid=test, batch_size=64, epochs=6, valid_freq=6, learning_rate=2e-3, logging_interval=1)
  assert model.is_fitted and model.dataset == train.name and "transcriptomic" in model.metadata
  loss = np.array(model.train_history["loss"])
  assert np.isfinite(loss).all()
  assert loss[-6:].mean() < loss[:6].mean()          # the ELBO improves
  assert len(model.valid_history["loss"]) >= 1
  # predict: every cell, in order; X batch_shape [S, N], Z batch_shape [N]
  pX, qZ = model.predict(test, sample_shape=3, batch_size=32, verbose=False)
  if name == "sisua":
    assert isinstance(pX, tuple) and pX[1].mean().shape == (3, len(test), 6)
    pX = pX[0]
  if name == "scvi":
    assert isinstance(qZ, tuple) and qZ[1].mean().shape == (len(test), 1)
    qZ = qZ[0]
  assert tuple(pX.batch_shape) == (3, len(test)) and tuple(pX.event_shape) == (60,)
  assert tuple(qZ.batch_shape) == (len(test),) and tuple(qZ.event_shape) == (10,)
  assert isinstance(pX.distribution, D.ZeroInflated)
  assert isinstance(pX.distribution.count_distribution, D.NegativeBinomialDisp)
  assert qZ.sample(1).shape == (1, len(test), 10)
  lp = pX.log_prob(torch.from_numpy(test.X).to(pX.mean().device))
  assert lp.shape == (3, len(test)) and torch.isfinite(lp).all()
  # log_prob of the returned distribution agrees with the per-cell llk the CUDA step computed itself
  llk_cuda = pX.elbo_terms[1].reshape(3, len(test))
  assert torch.allclose(lp, llk_cuda, rtol=2e-4, atol=1e-2)
  pX1, qZ1 = model.predict(test, batch_size=7, verbose=False)
  z_a = (qZ1[0] if isinstance(qZ1, tuple) else qZ1).mean().cpu().numpy()
  # save -> load -> identical latent means
  path = os.path.join(tmp_path, "model")
  model.save_weights(path)
  clone = load_model(path)
  assert type(clone) is type(model) and clone.dataset == model.dataset
  _, qZ2 = clone.predict(test, batch_size=64, verbose=False)
  z_b = (qZ2[0] if isinstance(qZ2, tuple) else qZ2).mean().cpu().numpy()
  np.testing.assert_allclose(z_a, z_b, rtol=1e-5, atol=1e-6)
  # posterior fast paths
  post = model.create_posterior(test, sample_shape=2, batch_size=16)
  assert post.imputed.shape == (len(test), 60) and post.latents.shape[0] == len(test)
  llk = post.cal_llk()
  assert np.isfinite(llk["original"]) and np.isfinite(llk["corrupted"])
  mll = model.marginal_log_prob(test.X[:8], library=torch.from_numpy(test.library[:8]).cuda() if name == "scvi" else None,
                                sample_shape=20) if name != "sisua" else None
  if mll is not None:
    assert mll.shape == (8,) and torch.isfinite(mll).all()
