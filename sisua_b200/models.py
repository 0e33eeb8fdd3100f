"""Host-side mirror of ``sisua.models`` for the ELBO train / infer hot path.

Same class names, constructor arguments, ``fit / predict / encode / decode / save_weights /
load_weights / create_posterior`` surface and error behaviour as the reference
(sisua/models/single_cell_model.py:67-306, vae.py:15-44, scvi.py:20-171, dca.py:13-28,
__init__.py:11-38); the arithmetic goes to libsisua_b200.so through ``Engine`` (ctypes).
Two HEAD defects are deliberately not copied: ``fit`` passing the undefined
name ``analytic`` and ``predict`` shuffling / truncating plain arrays."""
from __future__ import annotations

import os
import pickle
import time
import warnings
from typing import Any, Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import config as C
from . import distributions as D
from . import synthetic
from .config import NetConf, RVmeta
from .engine import Engine

__all__ = ["SingleCellModel", "VAE", "SISUA", "SCVI", "DeepCountAutoencoder", "NetConf", "RVmeta",
           "SingleCellData", "get_model", "get_all_models", "load_model"]


# ----------------------------------------------------------------------------------------------
# input contract (stands in for SingleCellOMIC.create_dataset, _single_cell_base.py:539-602)
# ----------------------------------------------------------------------------------------------
class SingleCellData:
  """Minimal stand-in for the slice of ``SingleCellOMIC`` the step consumes: a named count matrix,
  optional protein levels, variable names, dataset-level library statistics and the frozen
  semi-supervision mask of ``create_dataset(labels_percent=...)``."""

  def __init__(self, X, Y=None, name: str = "synthetic", var_names: Optional[Dict[str, Sequence[str]]] = None,
               labels_percent: float = 0.0, mask_seed: int = 1):
    self.X = np.ascontiguousarray(X, dtype=np.float32)
    self.Y = None if Y is None else np.ascontiguousarray(Y, dtype=np.float32)
    self.name = name
    self.var_names = var_names or {
        "transcriptomic": [f"gene{i}" for i in range(self.X.shape[1])],
        **({"proteomic": [f"prot{i}" for i in range(self.Y.shape[1])]} if self.Y is not None else {})}
    self.library = synthetic.library_stats(self.X)
    lp = 0.0 if self.Y is None else float(np.clip(labels_percent, 0.0, 1.0))   # forced to 0 for one OMIC (:578-579)
    self.mask = synthetic.label_mask(self.X.shape[0], lp, mask_seed) if lp > 0 else \
        np.zeros(self.X.shape[0], dtype=np.uint8)

  def __len__(self):
    return self.X.shape[0]

  @property
  def n_genes(self):
    return self.X.shape[1]

  def split(self, train_percent: float = 0.8, seed: int = 1):
    rng = np.random.RandomState(seed)
    idx = rng.permutation(len(self))
    n = int(train_percent * len(self))
    mk = lambda ids, tag: self._take(ids, f"{self.name}_{tag}")
    return mk(np.sort(idx[:n]), "train"), mk(np.sort(idx[n:]), "test")

  def _take(self, ids, name, recompute_library=True):
    """Row subset.  The library statistics are those of the SUBSET, as in the reference (create_dataset recomputes them on
    whatever dataset it is called on: _single_cell_base.py:566-570); data-parallel shards of one training set pass
    recompute_library=False so that every rank keeps the statistics of the whole training set."""
    out = SingleCellData.__new__(SingleCellData)
    out.X = self.X[ids]; out.Y = None if self.Y is None else self.Y[ids]
    out.name = name; out.var_names = self.var_names
    out.library = synthetic.library_stats(out.X) if recompute_library else self.library[ids]
    out.mask = self.mask[ids]
    return out


def _is_batch_iterable(x) -> bool:
  """True for an iterable of minibatches (a generator, a tf.data-like object, a list of batch dicts) as opposed to a whole
  dataset (SingleCellData, an array, a (X, Y) pair, a dict of arrays)."""
  if isinstance(x, (SingleCellData, np.ndarray, torch.Tensor, dict)) or x is None:
    return False
  if isinstance(x, (tuple, list)):
    return len(x) > 0 and all(isinstance(b, dict) and "inputs" in b for b in x)
  return hasattr(x, "__iter__")


def _to_data(x, require_meta=False) -> SingleCellData:
  if isinstance(x, SingleCellData):
    return x
  if isinstance(x, dict):
    return SingleCellData(x["x"], x.get("y"), name=x.get("name", "array"))
  if isinstance(x, (tuple, list)):
    return SingleCellData(x[0], x[1] if len(x) > 1 else None, name="array")
  return SingleCellData(x, name="array")


# ----------------------------------------------------------------------------------------------
class _Posterior:
  """What callers read from ``model.posteriors[i]`` / ``model.output_layers[i]``."""

  def __init__(self, rv: RVmeta):
    self.name = rv.name
    self.event_shape = (rv.dim,)
    self.posterior = rv.posterior
    self.is_zero_inflated = rv.is_zero_inflated
    self.trainable = True


class SingleCellModel:
  r""" Note: seed the model (``seed=...``) for reproducible results. """

  _kind = C.MODEL_VAE

  def __init__(self,
               outputs: RVmeta,
               latents: RVmeta = None,
               encoder: NetConf = None,
               decoder: NetConf = None,
               log_norm=True,
               beta=1.0,
               name=None,
               **kwargs):
    latents = RVmeta(10, 'diag', True, 'Latents') if latents is None else latents
    encoder = NetConf([64, 64], batchnorm=True, input_dropout=0.3) if encoder is None else encoder
    decoder = NetConf([64, 64], batchnorm=True) if decoder is None else decoder
    outs = list(outputs) if isinstance(outputs, (list, tuple)) else [outputs]
    labels = kwargs.pop("labels", None)
    labels = [] if labels is None else (list(labels) if isinstance(labels, (list, tuple)) else [labels])
    self.init_args = dict(outputs=outputs, latents=latents, encoder=encoder, decoder=decoder, log_norm=log_norm,
                          beta=beta, name=name, **({"labels": labels} if labels else {}), **kwargs)
    self._outputs, self.labels, self._latents = outs, labels, latents
    self._encoder, self._decoder = encoder, decoder
    self._log_norm = bool(log_norm)
    self.beta = float(beta)
    self.alpha = float(kwargs.pop("alpha", 10.0))
    self.name = name or type(self).__name__
    self._device_index = int(kwargs.pop("device", torch.cuda.current_device() if torch.cuda.is_available() else 0))
    self._gemm_mode = int(kwargs.pop("gemm_mode", C.GEMM_TC_3XFP16))
    self._seed = int(kwargs.pop("seed", 8))
    self._max_batch = int(kwargs.pop("max_batch", 8192))
    self._cfg_overrides = {k: kwargs.pop(k) for k in list(kwargs) if k in (
        "mean_act", "disp_act", "scale_act", "scvi_reapply_act", "mask_norm", "clip_mode")}
    for k in ("reduce_latent", "input_shape", "step", "path", "analytic"):
      kwargs.pop(k, None)
    if kwargs:
      raise TypeError(f"unexpected arguments: {sorted(kwargs)}")
    self.dataset = None
    self.metadata: Dict[str, Any] = dict()
    self.posteriors = [_Posterior(rv) for rv in self._outputs + self.labels]
    self.output_layers = self.posteriors[:len(self._outputs)]
    self.train_history: Dict[str, List[float]] = {}
    self.valid_history: Dict[str, List[float]] = {}
    self.is_fitted = False
    self._engine: Optional[Engine] = None
    self._pending_weights = None
    self.step = 0

  # ---------------------------------------------------------------- configuration -> StepConfig
  def _step_config(self) -> C.StepConfig:
    rv = self._outputs[0]
    if rv.posterior not in ("zinbd", "nbd", "zinb", "nb"):
      raise ValueError(f"posterior '{rv.posterior}' is outside the B200 hot path (zinbd, nbd, zinb, nb)")
    enc, dec = self._encoder, self._decoder
    if len(set(enc.units + dec.units)) != 1:
      raise ValueError("all hidden layers must share one width")
    if bool(enc.batchnorm) != bool(dec.batchnorm):
      raise ValueError("encoder and decoder must agree on batchnorm (one StepConfig flag drives both stacks)")
    lat = self._latents.posterior
    if not (lat in ("diag", "mvndiag", "normal") or self._latents.is_deterministic):
      raise ValueError(f"latent posterior '{lat}' is outside the B200 hot path ('diag', or deterministic 'relu' / 'linear')")
    if self.labels and self._kind != C.MODEL_SISUA:
      raise ValueError("label heads are implemented for SISUA only (the hot path of vae / scvi / dca has no protein head)")
    kw = dict(self._cfg_overrides)
    for k in ("mean_act", "disp_act", "scale_act"):
      if k in rv.kwargs:
        kw.setdefault(k, rv.kwargs[k])
    n_prot, y_dist = 0, "nb"
    if self.labels:
      if len(self.labels) != 1 or self.labels[0].posterior not in ("nb", "nbd"):
        raise ValueError("the semi-supervised head on the hot path is one 'nb' / 'nbd' protein RV")
      n_prot, y_dist = self.labels[0].dim, self.labels[0].posterior
    extra = self._extra_config()
    return C.make_step_config(
        C.MODEL_NAMES[self._kind], n_genes=rv.dim, n_proteins=n_prot, n_latent=self._latents.dim,
        n_hidden=enc.units[0], n_enc_layers=len(enc.units), n_dec_layers=len(dec.units), batchnorm=enc.batchnorm,
        log_norm=self._log_norm, x_dist=rv.posterior, y_dist=y_dist, gemm_mode=self._gemm_mode,
        max_batch=self._max_batch, input_dropout=enc.input_dropout, enc_dropout=enc.dropout,
        dec_dropout=dec.dropout, beta=self.beta, alpha=self.alpha,
        latent_linear=self._latents.posterior in ("linear", "identity", "deterministic"), **extra, **kw)

  def _extra_config(self) -> Dict[str, Any]:
    return {}

  @property
  def engine(self) -> Engine:
    if self._engine is None:
      self._engine = Engine(self._step_config(), self._device_index, seed=self._seed)
      if self._pending_weights is not None:
        self._restore(self._pending_weights)
        self._pending_weights = None
    return self._engine

  # ---------------------------------------------------------------- reference attribute surface
  def set_metadata(self, sco):
    if not isinstance(sco, SingleCellData):
      raise AssertionError(f"sco must be instance of SingleCellData but given: {type(sco)}")
    self.dataset = sco.name
    for k, v in sco.var_names.items():
      self.metadata[k] = list(v)
    return self

  @property
  def log_norm(self):
    return self._log_norm

  @property
  def is_zero_inflated(self):
    return self.posteriors[0].is_zero_inflated

  @property
  def is_semi_supervised(self):
    return len(self.labels) > 0

  @classmethod
  def _id(cls):
    return ''.join(c for c in cls.__name__ if c.isupper()).lower()

  class _ClassProp:
    def __get__(self, obj, owner):
      return owner._id()

  id = _ClassProp()

  # ---------------------------------------------------------------- one batch through the C ABI
  def _batch_tensors(self, data: SingleCellData, idx: torch.Tensor, dev_cache: Dict[str, torch.Tensor]):
    b = dict(x=dev_cache["x"][idx])
    if self.labels:
      b["y"] = dev_cache["y"][idx]
      b["mask"] = dev_cache["mask"][idx]
    if self._kind == C.MODEL_SCVI:
      b["library"] = dev_cache["library"][idx]
    return b

  def _upload(self, data: SingleCellData, storage: str = "float32") -> Dict[str, torch.Tensor]:
    dev = self.engine.device
    X = data.X
    compact = False
    if storage in ("uint16", "auto") and X.size:
      # counts are integers: uint16 holds them exactly in half the HBM (sisua_train_step_gather_u16 widens them in-kernel)
      compact = float(X.min()) >= 0 and float(X.max()) < 65536 and all(
          np.array_equal(X[s:s + 65536], np.rint(X[s:s + 65536])) for s in range(0, X.shape[0], 65536))
      if storage == "uint16" and not compact:
        raise ValueError("storage='uint16' needs non-negative integer counts below 65 536")
    if compact:
      cache = dict(x=torch.from_numpy(X.astype(np.uint16).view(np.int16)).to(dev))
    else:
      cache = dict(x=torch.from_numpy(X).to(dev))
    self.engine.set_count_bound(float(data.X.max()) if data.X.size else 0.0)     # keeps the fp16 gradient tiles in range
    if self.labels:
      if data.Y is None:
        raise ValueError("semi-supervised model needs the protein matrix Y")
      cache["y"] = torch.from_numpy(data.Y).to(dev)
      cache["mask"] = torch.from_numpy(data.mask).to(dev)
    if self._kind == C.MODEL_SCVI:
      cache["library"] = torch.from_numpy(data.library).to(dev)
    return cache

  def __call__(self, inputs, library=None, mask=None, training=False, sample_shape=(), eps=None, **kwargs):
    """One batch -> (pX_Z, qZ_X) like ``self(**data, training=..., sample_shape=S)`` (single_cell_model.py:178).
    ``training=True`` evaluates the training-mode graph (batch statistics in BatchNorm, dropout, moving statistics
    updated) without touching the weights: the parameters it returns are those of a forward pass of ``fit``."""
    eng = self.engine
    xs = list(inputs) if isinstance(inputs, (list, tuple)) else [inputs]
    x = eng._dev(xs[0])
    y = eng._dev(xs[1]) if len(xs) > 1 and self.labels else None
    B = x.shape[0]
    S = int(np.prod(sample_shape)) if sample_shape not in ((), None, 0) else None
    if self._kind == C.MODEL_SCVI and library is None:
      raise ValueError("scVI needs the `library` [B,2] statistics of the batch")
    if self.labels and y is None:
      y = torch.zeros((B, eng.cfg.n_proteins), device=eng.device)
    eps = eps or {}
    if training:
      if S not in (None, 1):
        raise ValueError("training-mode calls use sample_shape=() (configs/base.yaml:53)")
      self._train_calls = getattr(self, "_train_calls", 0) + 1
      eng.train_forward(x, y=y, library=library, mask=mask, seed=self._seed, step=self._train_calls, **eps)
      out = eng.last_forward_outputs(B)
      return self._wrap(out, B, None)
    if not eps:
      eng.set_infer_seed(self._seed, self.step)
    out = eng.infer(x, y=y, library=library, mask=mask, S=S or 1, want_mean=True, want_disp=True, want_pi=True, **eps)
    return self._wrap(out, B, S)

  def _wrap(self, out, B, S):
    cfg = self.engine.cfg
    G = cfg.n_genes
    shp = (lambda t, n: t.reshape(S, B, n)) if S else (lambda t, n: t.reshape(B, n))
    nb = D.NegativeBinomialDisp(shp(out["mean"], G), shp(out["disp"], G))
    base = D.ZeroInflated(nb, shp(out["pi_logit"], G)) if cfg.n_out_heads == 3 else nb
    pX = D.Independent(base, 1, name=self.posteriors[0].name)
    pX.elbo_terms = out["terms"]
    qZ = self._wrap_latents(out)
    if self.labels:
      pY = D.Independent(D.MeanOnly(shp(out["y_mean"], cfg.n_proteins)), 1, name=self.posteriors[1].name)
      return (pX, pY), qZ
    return pX, qZ

  def _wrap_latents(self, out):
    if self._kind == C.MODEL_DCA:
      qZ = D.VectorDeterministic(out["z_loc"], name=self._latents.name)
    else:
      qZ = D.MultivariateNormalDiag(out["z_loc"], out["z_scale"], name=self._latents.name)
    if self._kind == C.MODEL_SCVI:
      qL = D.Independent(D.Normal(out["lib_loc"][:, None], out["lib_scale"][:, None]), 1, name="Library")
      qZ = (qZ, qL)
    return qZ

  def encode(self, inputs, library=None, training=None, mask=None, sample_shape=(), **kwargs):
    return self(inputs, library=library, mask=mask, training=bool(training), sample_shape=sample_shape, **kwargs)[1]

  def decode(self, latents=None, training=None, mask=None, sample_shape=(), inputs=None, library=None, **kwargs):
    r""" ``decode(latents)`` (single_cell_model.py:141-151): latent samples ``[..., z]`` (a tensor, or the distribution
    returned by ``encode``, which is sampled once) -> output distribution(s) through the decoder-only entry point
    ``sisua_decode``.  scVI takes ``latents=(z, log_library)``.  ``inputs=`` keeps the round-1 behaviour (encode then
    decode in one fused call). """
    if latents is None:
      if inputs is None:
        raise ValueError("decode() needs `latents` (or `inputs=` for the fused encode -> decode call)")
      return self(inputs, library=library, mask=mask, training=bool(training), sample_shape=sample_shape, **kwargs)[0]
    eng = self.engine
    lib_s = None
    if self._kind == C.MODEL_SCVI:
      if not isinstance(latents, (tuple, list)) or len(latents) != 2:
        raise ValueError("scVI decodes (z, log_library)")
      latents, lib_s = latents
      lib_s = lib_s.sample() if isinstance(lib_s, D.Distribution) else lib_s
    z = latents.sample() if isinstance(latents, D.Distribution) else latents
    z = eng._dev(z)
    lead = tuple(z.shape[:-1])
    z2 = z.reshape(-1, z.shape[-1]).contiguous()
    out = eng.decode(z2, None if lib_s is None else eng._dev(lib_s).reshape(-1).contiguous())
    cfg = eng.cfg
    G = cfg.n_genes
    rs = lambda t, n: t.reshape(lead + (n,))
    nb = D.NegativeBinomialDisp(rs(out["mean"], G), rs(out["disp"], G))
    base = D.ZeroInflated(nb, rs(out["pi_logit"], G)) if cfg.n_out_heads == 3 else nb
    pX = D.Independent(base, 1, name=self.posteriors[0].name)
    if self.labels:
      return pX, D.Independent(D.MeanOnly(rs(out["y_mean"], cfg.n_proteins)), 1, name=self.posteriors[1].name)
    return pX

  # ---------------------------------------------------------------- predict
  def predict(self, inputs, sample_shape=(), batch_size=32, verbose=True, device="GPU", seed=None):
    r""" Predict on minibatches then return a single distribution (single_cell_model.py:153-211).  Row order is preserved
    and every cell is kept.  ONE streamed pass computes the latent statistics and the per-cell ELBO terms; the output
    distribution is *streamed* (``sisua_b200.streamed``): its ``[S, N, G]`` parameters are never stored -- ``log_prob``
    and ``mean_over_samples`` run inside the fused kernels chunk by chunk, ``mean() / variance() / sample()`` build
    dense tensors only when called.  ``device='CPU'`` returns the (small) latent statistics on the host. """
    assert device in ("CPU", "GPU"), f"Only support device CPU or GPU, but given: {device}"
    from . import streamed as ST
    data = _to_data(inputs)
    eng = self.engine
    cache = self._upload(data)
    N = len(data)
    S = int(np.prod(sample_shape)) if sample_shape not in ((), None, 0) else None
    # minibatch size does not change inference results (moving-average BN): the largest chunk that fits is used
    rows_per_call = max(1, eng.cfg.max_batch // (S or 1))
    src = ST.StreamSource(eng, cache, N, S, rows_per_call, seed=(self._seed + 12345) if seed is None else int(seed))
    Z, cfg = eng.cfg.n_latent, eng.cfg
    f = lambda *shape: torch.empty(shape, dtype=torch.float32, device=eng.device)
    merged = dict(z_loc=f(N, Z), z_scale=f(N, Z), terms=f(5, S or 1, N))
    if self._kind == C.MODEL_SCVI:
      merged["lib_loc"], merged["lib_scale"] = f(N), f(N)
    y_mean = f(S or 1, N, cfg.n_proteins) if self.labels else None
    for k, sl in src.chunks():
      out = src.run(k, sl)
      n = sl.stop - sl.start
      merged["z_loc"][sl] = out["z_loc"]; merged["z_scale"][sl] = out["z_scale"]
      merged["terms"][:, :, sl] = out["terms"].reshape(5, S or 1, n)
      if self._kind == C.MODEL_SCVI:
        merged["lib_loc"][sl] = out["lib_loc"]; merged["lib_scale"][sl] = out["lib_scale"]
      if y_mean is not None:
        y_mean[:, sl] = out["y_mean"].reshape(S or 1, n, cfg.n_proteins)
    if device == "CPU":
      merged = {k: v.cpu() for k, v in merged.items()}
    base = ST.StreamedZeroInflated(src) if cfg.n_out_heads == 3 else ST.StreamedNB(src)
    pX = ST.StreamedIndependent(base, name=self.posteriors[0].name)
    pX.elbo_terms = merged["terms"]
    qZ = self._wrap_latents(merged)
    if self.labels:
      ym = y_mean if S else y_mean[0]
      pY = D.Independent(D.MeanOnly(ym.cpu() if device == "CPU" else ym), 1, name=self.posteriors[1].name)
      return (pX, pY), qZ
    return pX, qZ

  # ---------------------------------------------------------------- fit
  def fit(self,
          train,
          valid=None,
          metadata=None,
          batch_size=64,
          optimizer='adam',
          learning_rate=1e-3,
          clipnorm=100.,
          epochs=-1,
          max_iter=-1,
          valid_freq=500,
          sample_shape=(),
          checkpoint=None,
          earlystop_threshold=0.001,
          earlystop_patience=20,
          earlystop_min_epoch=-1,
          terminate_on_nan=True,
          logging_interval=2,
          shuffle=True,
          seed=None,
          verbose=False,
          cuda_graph='auto',
          data_on='auto',
          timing=None,
          **kwargs):
    r""" `Model.compile` + `Model.fit` of the reference in one call
    (single_cell_model.py:213-236; keys of configs/base.yaml:45-62). """
    if isinstance(train, SingleCellData):
      self.set_metadata(train)
    elif isinstance(valid, SingleCellData):
      self.set_metadata(valid)
    elif isinstance(metadata, SingleCellData):
      self.set_metadata(metadata)
    if self.dataset is None or len(self.metadata) == 0:
      raise RuntimeError("First time call `fit`, set the 'metadata' argument to a "
                         "SingleCellData dataset to keep the dataset name and OMICs' "
                         "variables description.")
    # the remaining keys of configs/base.yaml:45-62 are honoured or rejected, never silently swallowed
    dp_shard = bool(kwargs.pop("dp_shard", True))     # data parallel: False = `train` already is this rank's shard
    dp_exchange = str(kwargs.pop("dp_exchange", "peer"))   # 'peer': one kernel over NVLink peer memory; 'nccl': all-reduce + Adam
    storage = str(kwargs.pop("storage", "auto"))           # resident training shard: 'float32', 'uint16' (exact for counts) or 'auto'
    if storage not in ("float32", "uint16", "auto"):
      raise ValueError("storage must be 'float32', 'uint16' or 'auto'")
    if dp_exchange not in ("peer", "nccl"):
      raise ValueError("dp_exchange must be 'peer' or 'nccl'")
    valid_interval = float(kwargs.pop("valid_interval", 0) or 0)          # seconds between validation passes (0: valid_freq only)
    allow_rollback = bool(kwargs.pop("allow_rollback", False))            # restore the best validated weights when stopping
    if int(kwargs.pop("earlystop_progress_length", 0) or 0) != 0:
      raise NotImplementedError("earlystop_progress_length > 0 (progress-based early stopping) is not implemented")
    for k in ("log_tag", "log_path", "skip_fitted", "compile_graph", "track_gradient_norm"):     # host-side logging knobs
      if k in kwargs:
        warnings.warn(f"fit(): '{k}' only affects the reference's TensorBoard / tf.function plumbing and is ignored here")
        kwargs.pop(k)
    if kwargs:
      raise TypeError(f"fit() got unexpected arguments: {sorted(kwargs)}")
    if str(optimizer).lower() != 'adam':
      raise ValueError("the fused optimiser on the hot path is Adam (configs/base.yaml:46)")
    if sample_shape not in ((), None, [], 0, 1, (1,)):
      raise NotImplementedError("training uses sample_shape=() (configs/base.yaml:53)")
    if _is_batch_iterable(train):
      # an iterable of minibatch dicts -- what SingleCellOMIC.create_dataset / tf.data yields and odin-ai's fit consumes
      # (sisua/data/_single_cell_base.py:582-601): every batch goes through sisua_train_step + sisua_adam_step as it comes
      return self._fit_batches(train, epochs=epochs, max_iter=max_iter, lr=float(learning_rate), clipnorm=float(clipnorm or 0.0),
                               terminate_on_nan=terminate_on_nan, logging_interval=logging_interval, verbose=verbose)
    train = _to_data(train)
    valid = _to_data(valid) if valid is not None else None
    eng = self.engine
    # data parallel over cells when torch.distributed is initialised (one process per GPU): every rank trains on its
    # contiguous shard, gradients are all-reduced (sisua_b200/distributed.py), replicas stay bit-identical
    from . import distributed as DP
    import torch.distributed as dist
    world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
    rank = dist.get_rank() if world > 1 else 0
    px = None
    if world > 1 and dp_exchange == "peer":
      px = getattr(self, "_peer_exchange", None)
      if px is None:
        # moves params / grads into symmetric memory; if ANY rank cannot (no P2P between the GPUs, no symmetric memory),
        # all ranks fall back to the NCCL all-reduce together -- a rank on its own path would dead-lock the others
        err = None
        try:
          px = DP.PeerExchange(eng)
        except Exception as e:      # noqa: BLE001
          err, px = f"{type(e).__name__}: {e}", None
        ok = torch.tensor([0.0 if err else 1.0], device=eng.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if ok.item() < 1.0:
          warnings.warn(f"peer-memory gradient exchange unavailable ({err or 'on another rank'}); using the NCCL all-reduce")
          px = None
        self._peer_exchange = px
    if world > 1 and dp_shard:
      b0, e0 = DP.shard_range(len(train), rank, world)
      n_common = len(train) // world                      # every rank takes the same number of steps
      train = train._take(np.arange(b0, b0 + n_common), train.name, recompute_library=False)
    if world > 1:
      DP.broadcast_parameters(eng.params, eng.bn_moving)
    reducer = DP.OverlappedAllReduce(eng) if px is None else None
    B = int(batch_size)
    if B > eng.cfg.max_batch:
      raise ValueError(f"batch_size {B} > max_batch {eng.cfg.max_batch}")
    N = len(train)
    if N < B:
      raise ValueError("dataset smaller than one batch (drop_remainder=True, train.py:126-135)")
    steps_per_epoch = N // B
    if epochs is None or epochs <= 0:
      epochs = 1 if (max_iter is None or max_iter <= 0) else int(np.ceil(max_iter / steps_per_epoch))
    total = epochs * steps_per_epoch if (max_iter is None or max_iter <= 0) else min(int(max_iter), epochs * steps_per_epoch)
    if data_on not in ("auto", "device", "host"):
      raise ValueError("data_on must be 'auto', 'device' or 'host'")
    host_stream = data_on == "host"
    vcache = self._upload(valid) if valid is not None else None
    rng_seed = (self._seed if seed is None else int(seed)) + 7919 * rank
    gen = torch.Generator(device=eng.device); gen.manual_seed(rng_seed)
    terms = torch.empty((5, B), device=eng.device)
    loss = torch.empty((1,), device=eng.device)
    names = ["loss", "llk_" + self.posteriors[0].name] + (["llk_" + self.posteriors[1].name] if self.labels else []) + \
        ["kl_" + self._latents.name] + (["kl_Library"] if self._kind == C.MODEL_SCVI else [])
    for n in names:
      self.train_history.setdefault(n, [])
      self.valid_history.setdefault(n, [])
    lr, cn = float(learning_rate), float(clipnorm or 0.0)
    step_seed = self._seed + 7919 * rank          # dropout masks and reparameterisation noise: Philox(step_seed; ..., step, stream)
    graphed = pipe = None
    if host_stream:
      # the dataset stays in (pinned) host memory and every step ships its minibatch over PCIe: CSR for integer counts
      # (single-cell matrices are 70-96 % zeros), the library's host-buffer entry point otherwise
      from .pipeline import HostDataset, HostTrainPipeline
      # plain VAE / DCA steps replay a CUDA graph per slot: those take the 2-bytes-per-non-zero packed CSR form
      hds = HostDataset(train, B, shuffle=shuffle, seed=rng_seed, with_y=bool(self.labels), with_library=self._kind == C.MODEL_SCVI,
                        packed=(self._kind in (C.MODEL_VAE, C.MODEL_DCA) and not self.labels))
      eng.set_count_bound(hds.max_count)
      eng.reset_step_counter(self.step)
      # (4 slots instead of 2 did not help the 8-GPU run: 190 M vs 204 M cells/s end to end; SISUA_HOST_DEPTH for experiments)
      pipe = HostTrainPipeline(eng, B, depth=int(os.environ.get("SISUA_HOST_DEPTH", 2)))
      host_losses = []
    else:
      cache = self._upload(train, storage)
      # launch-bound regime (the reference's minibatch sizes): replay the whole step as one CUDA graph
      use_graph = (cuda_graph is True) or (cuda_graph == 'auto' and B <= 2048 and world == 1)
      if use_graph:
        from .pipeline import GraphedGatherStep
        eng.reset_step_counter(self.step)
        graphed = GraphedGatherStep(eng, B, cache["x"], y_all=cache.get("y"), library_all=cache.get("library"),
                                    mask_all=cache.get("mask"), lr=lr, clipnorm=cn, seed=step_seed)
        terms, loss = graphed.terms, graphed.loss
    eng.nonfinite(reset=True)
    ep_loss = [torch.empty((1,), dtype=torch.float32).pin_memory() for _ in range(2)] if host_stream else None
    ep_pending = None
    nan_slot = [torch.empty((1,), dtype=torch.float32).pin_memory() for _ in range(2)]
    nan_pending = None
    nan_events = [None, None]
    best, patience, done = float("inf"), 0, 0
    best_state, last_valid = None, time.perf_counter()
    log_buf: List[torch.Tensor] = []
    stop = False
    t_start = None
    for ep in range(epochs):
      if not host_stream:
        perm = (torch.randperm(N, device=eng.device, generator=gen) if shuffle else torch.arange(N, device=eng.device)).to(torch.int32)
      for s in range(steps_per_epoch):
        if done >= total:
          stop = True
          break
        if timing is not None and done == int(timing.get("skip", 0)):
          torch.cuda.current_stream(eng.device).synchronize()
          if world > 1:
            dist.barrier()
          t_start = (time.perf_counter(), done)
        self.step += 1
        if host_stream:
          xb, extras = hds.batch(ep, s)
          host_losses.append(pipe.step(xb, None, step=self.step, lr=lr, clipnorm=cn, world=world, peer=px,
                                       allreduce=(lambda g: dist.all_reduce(g)) if (world > 1 and px is None) else None,
                                       seed=step_seed, **extras))
        elif graphed is not None:
          graphed.step(perm[s * B:(s + 1) * B])
        else:
          # one minibatch = B row indices into the matrices resident in HBM; the kernels gather the count rows themselves
          # and draw the reparameterisation noise in-kernel (no per-step torch kernels on the hot path)
          eng.train_step_gather(cache["x"], perm[s * B:(s + 1) * B], y_all=cache.get("y"), library_all=cache.get("library"),
                                mask_all=cache.get("mask"), terms=terms, loss=loss, seed=step_seed, step=self.step)
          if px is not None:      # reduce-scatter + clip + sharded Adam + all-gather: one kernel over peer memory
            px.step(lr=lr, clipnorm=cn, t=self.step)
          else:
            gscale = reducer()
            eng.adam_step(lr=lr, clipnorm=cn, grad_scale=gscale, t=self.step)
        done += 1
        # per-step NaN watch (configs/base.yaml:59): the step kernels raise a word in mapped host memory, read here without
        # a synchronisation.  One rank leaving alone would dead-lock the others, so N > 1 decides at the epoch boundary.
        if terminate_on_nan and world == 1 and not host_stream and done % 8 == 0:
          # bound the host's run-ahead to 16 steps, so the word below is never older than that (an event wait on a step
          # that finished long ago costs nothing once the queue is full anyway)
          if nan_events[1] is not None:
            nan_events[1].synchronize()
          nan_events = [torch.cuda.Event(), nan_events[0]]
          nan_events[0].record(torch.cuda.current_stream(eng.device))
        if terminate_on_nan and world == 1 and eng.nonfinite():
          torch.cuda.current_stream(eng.device).synchronize()      # (steps already queued would raise the word again)
          eng.nonfinite(reset=True)
          raise FloatingPointError(f"training loss is not finite at or before step {self.step} (terminate_on_nan, configs/base.yaml:59)")
        if logging_interval and done % int(logging_interval) == 0 and not host_stream:
          log_buf.append(torch.cat([loss, terms[1:].mean(dim=1)]))
        due = valid_freq and done % int(valid_freq) == 0
        if valid is not None and valid_interval > 0 and time.perf_counter() - last_valid >= valid_interval:
          due = True
        if world > 1 and valid is not None and valid_interval > 0:       # wall clocks differ between ranks
          f = torch.tensor([1.0 if due else 0.0], device=eng.device); dist.all_reduce(f, op=dist.ReduceOp.MAX); due = bool(f.item() > 0)
        if valid is not None and due:
          last_valid = time.perf_counter()
          v = self._evaluate(valid, vcache, B)
          if world > 1:       # every rank must take the same early-stop / NaN decision (BatchNorm moving statistics differ)
            vt = torch.tensor(v, device=eng.device, dtype=torch.float64)
            dist.all_reduce(vt)
            v = (vt / world).tolist()
          self._log(self.valid_history, names, v)
          if terminate_on_nan and not np.isfinite(v[0]):
            raise FloatingPointError("validation loss is not finite")
          if v[0] < best - abs(earlystop_threshold) * abs(best if np.isfinite(best) else 1.0):
            best, patience = v[0], 0
            if allow_rollback:
              best_state = [t.clone() for t in (eng.params, eng.bn_moving, eng.adam_m, eng.adam_v)]
            if checkpoint is not None:
              checkpoint()
          else:
            patience += 1
            if earlystop_patience and patience >= earlystop_patience and ep >= earlystop_min_epoch:
              stop = True
              break
      if host_stream and host_losses:
        host_losses = []
        if pipe.last_loss_dev is not None:
          # epoch-end loss WITHOUT draining the pipeline: the last step's device loss is copied to a pinned slot behind
          # the queued work and read one epoch later (a full synchronisation here cost one step's latency per epoch --
          # 7-step epochs lost 8 % to it)
          slot = ep_loss[ep % 2]
          slot.copy_(pipe.last_loss_dev, non_blocking=True)
          ev = torch.cuda.Event(); ev.record(torch.cuda.current_stream(eng.device))
          harvested = []
          if ep_pending is not None:
            ep_pending[0].synchronize()
            harvested.append(float(ep_pending[1].item()))
          ep_pending = (ev, slot)
          vals = np.array(harvested, dtype=np.float64)[:, None] if harvested else None
        else:
          vals = np.array(pipe.flush_all([pipe.host_loss[(pipe.i - 1) % len(pipe.host_loss)]]), dtype=np.float64)[-1:, None]
        if vals is not None:
          for row in vals:
            self.train_history["loss"].append(float(row[0]))
        bad = vals is not None and not np.isfinite(vals).all()
      elif log_buf:
        vals = torch.stack(log_buf).cpu().numpy()
        log_buf = []
        for row in vals:
          self._log(self.train_history, names, self._select(row))
        bad = not np.isfinite(vals[:, 0]).all()
      else:
        vals, bad = None, False
      if world > 1 and terminate_on_nan:
        # collective decision (a rank leaving alone would dead-lock the others), taken one epoch late so that reading the
        # reduced flag never drains the pipeline: every rank sees the same flag at the same epoch
        flag = torch.tensor([1.0 if (bad or eng.nonfinite()) else 0.0], device=eng.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        nslot = nan_slot[ep % 2]
        nslot.copy_(flag, non_blocking=True)
        nev = torch.cuda.Event(); nev.record(torch.cuda.current_stream(eng.device))
        bad = False
        if nan_pending is not None:
          nan_pending[0].synchronize()
          bad = bool(nan_pending[1].item() > 0)
        nan_pending = (nev, nslot)
      if terminate_on_nan and bad:
        raise FloatingPointError("training loss is not finite (terminate_on_nan, configs/base.yaml:59)")
      if verbose and vals is not None:
        print(f"epoch {ep + 1}/{epochs} loss {vals[-1, 0]:.3f}")
      if stop:
        break
    if nan_pending is not None:         # the last epoch's collective flag
      nan_pending[0].synchronize()
      if nan_pending[1].item() > 0:
        raise FloatingPointError("training loss is not finite (terminate_on_nan, configs/base.yaml:59)")
    if ep_pending is not None:          # the last epoch's loss
      ep_pending[0].synchronize()
      self.train_history["loss"].append(float(ep_pending[1].item()))
      if terminate_on_nan and world == 1 and not np.isfinite(self.train_history["loss"][-1]):
        raise FloatingPointError("training loss is not finite (terminate_on_nan, configs/base.yaml:59)")
    if timing is not None and t_start is not None:
      torch.cuda.current_stream(eng.device).synchronize()
      if world > 1:
        dist.barrier()
      timing["seconds"] = time.perf_counter() - t_start[0]
      timing["steps"] = done - t_start[1]
      timing["cells_per_step"] = B * world
      if host_stream:
        timing["h2d_bytes_per_step"] = hds.h2d_bytes / max(1, hds.batches_served)
    if stop and allow_rollback and best_state is not None:
      for t, b_ in zip((eng.params, eng.bn_moving, eng.adam_m, eng.adam_v), best_state):
        t.copy_(b_)
    if world > 1:
      DP.average_moving_statistics(eng.bn_moving)
      if px is not None:
        px.gather_optimizer_state()      # the Adam moments were maintained per shard
    torch.cuda.current_stream(eng.device).synchronize()
    self.is_fitted = True
    return self

  def _select(self, row):
    # row = [loss, llk_x, llk_y, kl_z, kl_l] -> the metric names of this model
    out = [row[0], row[1]]
    if self.labels:
      out.append(row[2])
    out.append(row[3])
    if self._kind == C.MODEL_SCVI:
      out.append(row[4])
    return out

  @staticmethod
  def _log(hist, names, vals):
    for n, v in zip(names, vals):
      hist[n].append(float(v))

  def _evaluate(self, data: SingleCellData, cache, B):
    """Validation pass: mean ELBO terms over `data` (inference mode, one Philox sample per cell, fixed seed)."""
    eng = self.engine
    acc, cnt = torch.zeros(5, device=eng.device), 0
    for k, s in enumerate(range(0, len(data), B)):
      idx = torch.arange(s, min(len(data), s + B), device=eng.device)
      b = self._batch_tensors(data, idx, cache)
      eng.set_infer_seed(self._seed + 777, k)
      out = eng.infer(want_mean=False, **b)
      acc += out["terms"].sum(dim=1)
      cnt += idx.numel()
    m = (acc / cnt).cpu().numpy()
    return self._select(np.concatenate([[-m[0]], m[1:]]))

  def marginal_log_prob(self, inputs, library=None, mask=None, sample_shape=100, eps=None, **kwargs):
    """``(marginal, {output name: llk})`` as ``Posterior.cal_marginal_llk`` consumes it (sisua/analysis/posterior.py:964-968):
    the importance-weighted bound  log p(x) ~= logsumexp_s[log p(x|z_s) (+ alpha m log p(y|z_s)) + log p(z_s) - log q(z_s|x)]
    - log S  per cell, and per output  logsumexp_s log p(.|z_s) - log S.  All of it runs in ``sisua_marginal_llk``; a batch
    larger than max_batch / S cells is split."""
    S = int(np.prod(sample_shape)) if sample_shape not in ((), None, 0) else 1
    eng = self.engine
    xs = list(inputs) if isinstance(inputs, (list, tuple)) else [inputs]
    x = eng._dev(xs[0])
    y = eng._dev(xs[1]) if len(xs) > 1 and self.labels else None
    library = eng._dev(library); mask = eng._dev(mask, torch.uint8)
    B = x.shape[0]
    if S > eng.cfg.max_batch:
      raise ValueError(f"sample_shape {S} exceeds max_batch {eng.cfg.max_batch}")
    if self._kind == C.MODEL_SCVI and library is None:
      raise ValueError("scVI needs the `library` [B,2] statistics of the batch")
    if self.labels and y is None:
      y = torch.zeros((B, eng.cfg.n_proteins), device=eng.device)
    rows = max(1, eng.cfg.max_batch // S)
    m_parts, x_parts, y_parts = [], [], []
    eps = eps or {}
    for k, s0 in enumerate(range(0, B, rows)):
      sl = slice(s0, min(B, s0 + rows))
      if not eps:
        eng.set_infer_seed(self._seed + 4242, k)
      e = {n: v[:, sl] for n, v in eps.items()}
      m, lx, ly = eng.marginal_llk(x[sl], y=None if y is None else y[sl], library=None if library is None else library[sl],
                                   mask=None if mask is None else mask[sl], S=S, **e)
      m_parts.append(m); x_parts.append(lx)
      if ly is not None:
        y_parts.append(ly)
    llk = {self.posteriors[0].name: torch.cat(x_parts)}
    if y_parts:
      llk[self.posteriors[1].name] = torch.cat(y_parts)
    return torch.cat(m_parts), llk

  # ---------------------------------------------------------------- posterior / persistence
  def create_posterior(self, test_sco=None, dropout_rate=0.2, retain_rate=0.2, corrupt_distribution='binomial',
                       batch_size=8, sample_shape=10, reduce_latents=None, verbose=True, train_percent=0.8,
                       random_state=1, corrupt_on='host'):
    r""" Create a `Posterior` object for evaluation (single_cell_model.py:247-281); ``corrupt_on='device'`` corrupts the
    test set on the GPU (``sisua_corrupt_counts``) instead of with the reference's host routine. """
    if not self.is_fitted:
      raise RuntimeError("fit() must be called before creating Posterior.")
    if isinstance(test_sco, SingleCellData):
      test = test_sco
    elif self.dataset is None:
      raise ValueError("Call SingleCellModel.set_metadata() to track the fitted dataset.")
    else:
      raise ValueError(f"dataset '{self.dataset}' cannot be re-loaded here (no dataset registry on the hot path); "
                       "pass test_sco=")
    from .posterior import Posterior
    return Posterior(self, test, dropout_rate=dropout_rate, retain_rate=retain_rate,
                     corrupt_distribution=corrupt_distribution, batch_size=batch_size, sample_shape=sample_shape,
                     random_state=random_state, name=f"{self.id}_{self.dataset}", corrupt_on=corrupt_on)

  def _snapshot(self):
    eng = self.engine
    return dict(params=eng.params.cpu().numpy(), bn_moving=eng.bn_moving.cpu().numpy(),
                adam_m=eng.adam_m.cpu().numpy(), adam_v=eng.adam_v.cpu().numpy(), step=self.step,
                layout=[(e.name, e.offset, e.shape, e.ld) for e in eng.entries])

  def _restore(self, snap):
    eng = self._engine
    if [(e.name, e.offset, e.shape, e.ld) for e in eng.entries] != [tuple(x) for x in snap["layout"]]:
      raise ValueError("checkpoint layout does not match this model")
    eng.params.copy_(torch.from_numpy(snap["params"]))
    eng.bn_moving.copy_(torch.from_numpy(snap["bn_moving"]))
    eng.adam_m.copy_(torch.from_numpy(snap["adam_m"]))
    eng.adam_v.copy_(torch.from_numpy(snap["adam_v"]))
    self.step = int(snap["step"])

  def save_weights(self, filepath, overwrite=True):
    r""" weights blob + ``.metamodel`` pickle ``[class_name, dataset, metadata, init_args]``
    (single_cell_model.py:295-306); Adam state and BN moving statistics are persisted too. """
    if os.path.exists(filepath) and not overwrite:
      raise FileExistsError(filepath)
    snap = self._snapshot()
    # plain arrays + a JSON layout (no pickle: loading a weights file must not be able to execute code; only the
    # .metamodel sidecar is a pickle, as in the reference)
    import json
    with open(filepath, "wb") as f:
      np.savez(f, params=snap["params"], bn_moving=snap["bn_moving"], adam_m=snap["adam_m"], adam_v=snap["adam_v"],
               step=np.int64(snap["step"]), layout=np.frombuffer(json.dumps([list(x) for x in snap["layout"]]).encode(), dtype=np.uint8))
    with open(f"{filepath}.metamodel", "wb") as f:
      pickle.dump([self.__class__.__name__, self.dataset, self.metadata, dict(self.init_args)], f)
    return self

  def load_weights(self, filepath, raise_notfound=False, verbose=False):
    r""" Load all the saved weights at given path (single_cell_model.py:283-293) """
    if not os.path.exists(filepath):
      if raise_notfound:
        raise FileNotFoundError(filepath)
      return self
    import json
    with np.load(filepath, allow_pickle=False) as z:
      snap = dict(params=z["params"], bn_moving=z["bn_moving"], adam_m=z["adam_m"], adam_v=z["adam_v"], step=int(z["step"]),
                  layout=[(n, o, tuple(sh), ld) for n, o, sh, ld in json.loads(bytes(z["layout"]).decode())])
    if self._engine is None and not torch.cuda.is_available():
      self._pending_weights = snap
    else:
      _ = self.engine
      self._restore(snap)
    meta = f"{filepath}.metamodel"
    if os.path.exists(meta):
      with open(meta, "rb") as f:
        class_name, dataset, metadata, _ = pickle.load(f)
      assert class_name == self.__class__.__name__
      self.dataset, self.metadata = dataset, metadata
    self.is_fitted = True
    return self

  def _fit_batches(self, batches, epochs, max_iter, lr, clipnorm, terminate_on_nan, logging_interval, verbose):
    eng = self.engine
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
      raise NotImplementedError("fit(iterable of batches) is single-GPU; shard a SingleCellData over ranks instead")
    names = ["loss", "llk_" + self.posteriors[0].name] + (["llk_" + self.posteriors[1].name] if self.labels else []) + \
        ["kl_" + self._latents.name] + (["kl_Library"] if self._kind == C.MODEL_SCVI else [])
    for n in names:
      self.train_history.setdefault(n, [])
      self.valid_history.setdefault(n, [])
    eng.nonfinite(reset=True)
    eng.reset_step_counter(self.step)
    limit = int(max_iter) if (max_iter is not None and max_iter > 0) else None
    done, log_buf = 0, []
    for ep in range(max(1, int(epochs or 1))):
      seen = 0
      for b in batches:
        if limit is not None and done >= limit:
          break
        inputs = b["inputs"] if isinstance(b, dict) else b
        x, y = (tuple(inputs) + (None,))[:2] if isinstance(inputs, (tuple, list)) else (inputs, None)
        extra = b if isinstance(b, dict) else {}
        if self.labels and y is None:
          raise ValueError("this model has a label head: every batch needs inputs=(x, y)")
        if self._kind == C.MODEL_SCVI and extra.get("library") is None:
          raise ValueError("scVI batches need 'library' ([B, 2]: mean and variance of the log library size)")
        self.step += 1
        terms, loss = eng.train_step(x, y=y if self.labels else None, library=extra.get("library") if self._kind == C.MODEL_SCVI else None,
                                     mask=extra.get("mask") if self.labels else None, seed=self._seed, step=self.step)
        eng.adam_step(lr=lr, clipnorm=clipnorm, t=self.step)
        done += 1; seen += 1
        if terminate_on_nan and eng.nonfinite():
          torch.cuda.current_stream(eng.device).synchronize()
          eng.nonfinite(reset=True)
          raise FloatingPointError(f"training loss is not finite at or before step {self.step} (terminate_on_nan, configs/base.yaml:59)")
        if logging_interval and done % int(logging_interval) == 0:
          log_buf.append(torch.cat([loss, terms[1:].mean(dim=1)]))
      if log_buf:
        for row in torch.stack(log_buf).cpu().numpy():
          self._log(self.train_history, names, self._select(row))
        if verbose:
          print(f"epoch {ep + 1} loss {self.train_history['loss'][-1]:.3f}")
        log_buf = []
      if seen == 0 or (limit is not None and done >= limit):      # exhausted generator / step budget reached
        break
    self.is_fitted = True
    return self

  def export_keras(self, filepath, name_map=None):
    r""" Writes the weights as an ``.npz`` keyed by Keras variable names with Keras tensor conventions
    (``Dense.kernel`` = [in, out], BatchNormalization ``gamma / beta / moving_mean / moving_variance``, ``Adam/.../m|v``,
    ``Adam/iter``): what ``{v.name: v.numpy() for v in keras_model.variables}`` holds on the reference's side
    (sisua_b200/keras_layout.py; the reference's TF checkpoints, single_cell_model.py:295-306, need TensorFlow). """
    from . import keras_layout as KL
    snap = self._snapshot()
    arrays = KL.to_keras(self.engine.cfg, snap["params"], snap["bn_moving"], snap["adam_m"], snap["adam_v"], step=snap["step"],
                         name_map=name_map)
    with open(filepath, "wb") as f:
      np.savez(f, **arrays)
    return self

  def import_keras(self, filepath_or_arrays, name_map=None, strict=True):
    r""" Loads Keras-named arrays (an ``.npz`` path or a dict) into the model: kernels are transposed into the K-major
    flat layout, Adam slots and the step counter are taken if present. """
    from . import keras_layout as KL
    if isinstance(filepath_or_arrays, (str, os.PathLike)):
      with np.load(filepath_or_arrays, allow_pickle=False) as z:
        arrays = {k: z[k] for k in z.files}
    else:
      arrays = dict(filepath_or_arrays)
    eng = self.engine
    flat, moving, m, v, step = KL.from_keras(eng.cfg, arrays, name_map=name_map, strict=strict)
    eng.params.copy_(torch.from_numpy(flat))
    eng.bn_moving.copy_(torch.from_numpy(moving).reshape(eng.bn_moving.shape))
    if m is not None and v is not None:
      eng.adam_m.copy_(torch.from_numpy(m)); eng.adam_v.copy_(torch.from_numpy(v))
    if step is not None:
      self.step = int(step)
    self.is_fitted = True
    return self

  def plot_learning_curves(self, path=None, **kwargs):
    r""" Learning curves of ``fit`` (sisua/analysis/posterior.py:677-683 reads this off the model).  Draws with matplotlib
    when it is installed; otherwise writes the curves as CSV (``path`` required) -- the data, not the picture, is what
    the hot path produces. """
    hist = {"train_" + k: v for k, v in self.train_history.items()}
    hist.update({"valid_" + k: v for k, v in self.valid_history.items()})
    try:
      import matplotlib
      matplotlib.use("Agg")
      from matplotlib import pyplot as plt
    except ImportError:
      if path is None:
        raise RuntimeError("matplotlib is not installed: pass `path=` to get the learning curves as CSV")
      n = max((len(v) for v in hist.values()), default=0)
      with open(path, "w") as f:
        f.write(",".join(hist) + "\n")
        for i in range(n):
          f.write(",".join(str(v[i]) if i < len(v) else "" for v in hist.values()) + "\n")
      return path
    fig, axes = plt.subplots(1, max(1, len(self.train_history)), figsize=(4 * max(1, len(self.train_history)), 3))
    for ax, k in zip(np.atleast_1d(axes), self.train_history):
      ax.plot(self.train_history[k], label="train")
      if self.valid_history.get(k):
        ax.plot(np.linspace(0, len(self.train_history[k]), len(self.valid_history[k])), self.valid_history[k], label="valid")
      ax.set_title(k); ax.legend()
    if path is not None:
      fig.savefig(path)
    return fig


class VAE(SingleCellModel):
  r""" Variational Auto Encoder (sisua/models/vae.py:15-16) """
  _kind = C.MODEL_VAE


class SISUA(SingleCellModel):
  r""" Multi-task SemI-SUpervised Autoencoder (sisua/models/vae.py:19-44):
  transcriptomic ZINB/NB output + a protein head read from the same decoder output. """
  _kind = C.MODEL_SISUA

  def __init__(self, outputs, labels, **kwargs):
    super().__init__(outputs=outputs, labels=labels, **kwargs)
    if not self.labels:
      raise ValueError("SISUA needs `labels`")


class SCVI(SingleCellModel):
  r""" scVI re-implementation (sisua/models/scvi.py:20-171): library-size latent, softmax gene scale. """
  _kind = C.MODEL_SCVI

  def __init__(self, outputs, latents=None, library=None, encoder=None, encoder_l=None, clip_library=1e3, **kwargs):
    latents = RVmeta(10, 'diag', True, "Latents") if latents is None else latents
    self._library = RVmeta(1, 'normal', True, "Library") if library is None else library
    encoder = NetConf([64, 64], batchnorm=True, dropout=0.1, name='Encoder') if encoder is None else encoder
    self._encoder_l = NetConf([64], batchnorm=True, dropout=0.1, name='EncoderL') if encoder_l is None else encoder_l
    out0 = outputs[0] if isinstance(outputs, (list, tuple)) else outputs
    assert out0.posterior in ('zinbd', 'nbd'), \
        "scVI only support transcriptomic distribution: 'zinbd' or 'nbd', but given: %s" % str(outputs)
    self.clip_library = float(clip_library)
    if set(self._encoder_l.units) != set(encoder.units[:1]) or bool(self._encoder_l.batchnorm) != bool(encoder.batchnorm):
      raise ValueError("encoder_l must use the encoder's layer width and batchnorm setting (both first layers are one "
                       "fused [2H, G] tensor-core operand)")
    super().__init__(outputs, latents=latents, encoder=encoder, **kwargs)
    self.init_args.update(library=self._library, encoder_l=self._encoder_l, clip_library=clip_library)

  def _extra_config(self):
    return dict(n_encl_layers=len(self._encoder_l.units), encl_dropout=self._encoder_l.dropout,
                clip_library=self.clip_library)


class DeepCountAutoencoder(SingleCellModel):
  r""" Deep Count Autoencoder (sisua/models/dca.py:13-28): deterministic latent, no KL. """
  _kind = C.MODEL_DCA

  def __init__(self, outputs, latents=None, **kwargs):
    latents = RVmeta(10, 'relu', True, name="Latents") if latents is None else latents
    if not latents.is_deterministic:
      warnings.warn("DeepCountAutoencoder only support deterministic latents, "
                    f"but given {latents}, use default linear Dense layer for latents.")
      latents = latents.copy(posterior='linear')      # sisua/models/dca.py:22-27
    super().__init__(outputs=outputs, latents=latents, **kwargs)


# ----------------------------------------------------------------------------------------------
def get_all_models() -> list:
  found = [v for v in globals().values() if isinstance(v, type) and issubclass(v, SingleCellModel)]
  return sorted(found, key=lambda cls: cls.id)


def get_model(model):
  if isinstance(model, type):
    model = model.__name__
  model = str(model).lower()
  for key, val in list(globals().items()):
    if isinstance(val, type) and issubclass(val, SingleCellModel):
      if model == key.lower() or model == val.id:
        return val
  raise RuntimeError(f"Cannot find SingleCellModel with type '{model}'")


def load_model(filepath: str) -> SingleCellModel:
  with open(f"{filepath}.metamodel", 'rb') as f:
    class_name, dataset, metadata, kwargs = pickle.load(f)
  model = get_model(class_name)(**kwargs)
  model.load_weights(filepath, raise_notfound=True)
  return model
