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
  assert len(llk) == 4 and all(np.isfinite(v) for v in llk.values())
  lib = torch.from_numpy(test.library[:8]).cuda() if name == "scvi" else None
  inputs = (test.X[:8], test.Y[:8]) if name == "sisua" else test.X[:8]
  mll, parts = model.marginal_log_prob(inputs, library=lib, mask=test.mask[:8] if name == "sisua" else None, sample_shape=20)
  assert mll.shape == (8,) and torch.isfinite(mll).all() and "transcriptomic" in parts and parts["transcriptomic"].shape == (8,)
  # decode(latents): decoder-only entry point
  if gemm_mode == 1:
    z = torch.randn(5, 10, device="cuda")
    lat = (z, torch.full((5,), 6.0, device="cuda")) if name == "scvi" else z
    pz = model.decode(lat)
    pz = pz[0] if isinstance(pz, tuple) else pz
    assert pz.mean().shape == (5, 60) and torch.isfinite(pz.mean()).all()


@pytest.mark.gpu
def test_fit_argument_handling_and_host_streaming(tmp_path):
  """The remaining keys of configs/base.yaml:45-62 are honoured or rejected; fit(data_on='host') streams minibatches from
  pinned host memory (CSR) through HostTrainPipeline and trains like the device-resident path."""
  sco = _data(n=1024)
  train, valid = sco.split(0.75)
  m = _make("vae", seed=2)
  with pytest.raises(TypeError):
    m.fit(train, batch_size=64, epochs=1, no_such_option=1)
  with pytest.raises(NotImplementedError):
    m.fit(train, batch_size=64, epochs=1, earlystop_progress_length=5)
  timing = {"skip": 2}
  m.fit(train, valid=valid, batch_size=64, epochs=4, valid_freq=6, learning_rate=2e-3, logging_interval=1, allow_rollback=True,
        valid_interval=0, data_on="host", timing=timing)
  loss = np.array(m.train_history["loss"])
  assert np.isfinite(loss).all() and len(loss) >= 4
  assert timing["steps"] == 4 * (len(train) // 64) - 2 and timing["seconds"] > 0 and timing["h2d_bytes_per_step"] > 0
  m2 = _make("vae", seed=2)
  m2.fit(train, batch_size=64, epochs=4, learning_rate=2e-3, logging_interval=1)
  l2 = np.array(m2.train_history["loss"])
  assert l2[-6:].mean() < l2[:6].mean()
  # both paths reach a comparable loss (different minibatch orders, same data / model / optimiser)
  assert abs(loss[-1] - l2[-6:].mean()) < 0.25 * abs(l2[-6:].mean())
  p = m2.plot_learning_curves(path=os.path.join(tmp_path, "curves.csv"))
  assert p is not None


@pytest.mark.gpu
def test_terminate_on_nan_acts_within_the_epoch():
  """A non-finite training loss raises a word in mapped host memory from inside the step (sisua_nonfinite_flag); fit()
  polls it every step, so terminate_on_nan stops the run long before the end of the epoch (configs/base.yaml:59)."""
  sco = _data(n=2048)
  m = _make("vae", seed=3)
  m.fit(sco, batch_size=64, epochs=1, max_iter=2)          # builds the engine
  eng = m.engine
  assert not eng.nonfinite()
  e = [e for e in eng.entries if e.name == "out.b"][0]
  eng.params[e.offset:e.offset + 4].fill_(float("nan"))    # poison output-head biases (ReLU / clamps would launder a NaN further up)
  with pytest.raises(FloatingPointError):
    m.fit(sco, batch_size=64, epochs=1, terminate_on_nan=True)
  assert m.step < 2 + 2048 // 64, m.step                   # left before the epoch's 32 steps were issued
  assert not eng.nonfinite()                               # fit() cleared the word when it raised


@pytest.mark.gpu
def test_keras_named_weight_exchange(tmp_path):
  """export_keras / import_keras: a model rebuilt from the Keras-named arrays predicts exactly like the original."""
  sco = _data(n=512)
  m = _make("vae", seed=4)
  m.fit(sco, batch_size=64, epochs=1)
  path = os.path.join(tmp_path, "weights_keras.npz")
  m.export_keras(path)
  with np.load(path) as z:
    assert z["encoder/dense/kernel:0"].shape == (60, 64) and "Adam/iter:0" in z.files
  m2 = _make("vae", seed=99)
  m2.metadata, m2.dataset = m.metadata, m.dataset
  m2.import_keras(path)
  assert m2.step == m.step
  z1 = m.encode(sco.X[:32])
  z2 = m2.encode(sco.X[:32])
  a, b = (z1[0] if isinstance(z1, tuple) else z1), (z2[0] if isinstance(z2, tuple) else z2)
  assert torch.equal(a.mean(), b.mean())


@pytest.mark.gpu
def test_fit_accepts_an_iterable_of_batch_dicts():
  """fit(iterable of dict(inputs, library, mask)) -- the minibatch dicts SingleCellOMIC.create_dataset yields
  (sisua/data/_single_cell_base.py:582-601) -- trains batch by batch; a generator is consumed once."""
  sco = _data(n=640, p=6, labels_percent=0.5)
  m = _make("sisua", seed=5)
  m.set_metadata(sco)

  def batches():
    for s in range(0, 640, 64):
      yield dict(inputs=(sco.X[s:s + 64], sco.Y[s:s + 64]), mask=sco.mask[s:s + 64])

  m.fit(list(batches()), epochs=3, learning_rate=2e-3, logging_interval=1)
  loss = np.array(m.train_history["loss"])
  assert len(loss) == 30 and np.isfinite(loss).all() and loss[-5:].mean() < loss[:5].mean()
  assert m.step == 30 and m.is_fitted
  m.fit(batches(), epochs=3, logging_interval=1)           # a generator: one pass, then it is exhausted
  assert m.step == 40
  v = _make("scvi", seed=5); v.set_metadata(sco)
  with pytest.raises(ValueError):
    v.fit([dict(inputs=sco.X[:64])], epochs=1)              # scVI needs the library statistics with every batch


@pytest.mark.gpu
def test_fit_with_uint16_resident_storage():
  """fit(storage='uint16'): the training shard lives in HBM as uint16 and trains exactly like the float32 shard (same
  seeds -> same minibatches, masks and noise); non-integer data is refused."""
  sco = _data(n=1024, g=64)
  a = _make("vae", g=64, seed=6); b = _make("vae", g=64, seed=6)
  a.fit(sco, batch_size=128, epochs=3, learning_rate=2e-3, logging_interval=1, storage="uint16")
  b.fit(sco, batch_size=128, epochs=3, learning_rate=2e-3, logging_interval=1, storage="float32")
  la, lb = np.array(a.train_history["loss"]), np.array(b.train_history["loss"])
  assert len(la) == 24 and np.isfinite(la).all()
  np.testing.assert_allclose(la, lb, rtol=2e-3)
  frac = SingleCellData(sco.X + 0.5, name="frac")
  with pytest.raises(ValueError):
    _make("vae", g=64, seed=6).fit(frac, batch_size=128, epochs=1, storage="uint16")
  _make("vae", g=64, seed=6).fit(frac, batch_size=128, epochs=1, storage="auto")      # falls back to float32
