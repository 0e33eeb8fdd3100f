"""Streamed (lazy) output distributions of ``SingleCellModel.predict``.

The reference's ``predict`` (sisua/models/single_cell_model.py:153-211) concatenates the parameter tensors of every
minibatch, i.e. it materialises ``[S, N, 3 G]`` floats on the host; for the 1 M-cell x 2 000-gene scalability config
with the Posterior's 10 samples that is 240 GB.  Here ``predict`` makes ONE streamed pass (latent statistics and
per-cell ELBO terms, a few floats per cell) and hands back distribution objects that keep the counts in HBM and
RECOMPUTE their parameters chunk by chunk through ``sisua_infer_ex`` when a caller asks for something:

* ``log_prob(x)`` -> ``[S, N]`` straight from the fused likelihood epilogue (the ``[S, N, G]`` parameters never exist),
* ``mean_over_samples()`` -> ``[N, G]`` (the Posterior's "imputed" matrix, posterior.py:986-988) accumulated in the
  epilogue,
* ``mean() / variance() / sample()`` -> dense tensors, because that is what was asked for.

The reparameterisation noise is Philox(seed; row, column, call index, sample): a chunk always sees the same noise, so
every recomputation describes the same ``S`` posterior samples.  Class layout follows the slice of TFP the reference's
callers use (``Independent(ZeroInflated(NegativeBinomialDisp))``, ``.distribution.count_distribution``)."""
from __future__ import annotations

from typing import Dict, Iterator, Optional, Tuple

import torch

from . import distributions as D


class StreamSource:
  """What a streamed distribution needs to recompute a chunk: the engine, the device-resident inputs, the chunking and
  the noise seed."""

  def __init__(self, engine, cache: Dict[str, torch.Tensor], n_cells: int, S: Optional[int], rows_per_call: int, seed: int):
    self.eng, self.cache, self.N, self.S, self.rows, self.seed = engine, cache, int(n_cells), S, int(rows_per_call), int(seed)

  @property
  def samples(self) -> int:
    return self.S or 1

  @property
  def fused(self) -> bool:
    return self.eng.cfg.gemm_mode != 0 and not (self.eng.cfg.model_kind == 1 and self.eng.cfg.scvi_reapply_act)

  def chunks(self) -> Iterator[Tuple[int, slice]]:
    for k, s in enumerate(range(0, self.N, self.rows)):
      yield k, slice(s, min(self.N, s + self.rows))

  def run(self, k: int, sl: slice, x_eval: Optional[torch.Tensor] = None, **want) -> Dict[str, torch.Tensor]:
    c = self.cache
    self.eng.set_infer_seed(self.seed, k)            # chunk k always draws the same noise
    return self.eng.infer_ex(c["x"][sl], x_eval=None if x_eval is None else x_eval[sl], y=c["y"][sl] if "y" in c else None,
                             library=c["library"][sl] if "library" in c else None, mask=c["mask"][sl] if "mask" in c else None,
                             S=self.samples, **want)

  def shape(self, width: int) -> Tuple[int, ...]:
    return ((self.S, self.N, width) if self.S else (self.N, width))

  def gather(self, key: str, width: int, **want) -> torch.Tensor:
    """Dense [S, N, width] (or [N, width]) tensor of one parameter, chunk by chunk."""
    parts = []
    for k, sl in self.chunks():
      t = self.run(k, sl, want_latent=False, **want)[key]
      n = sl.stop - sl.start
      parts.append(t.reshape(self.samples, n, width))
    full = torch.cat(parts, dim=1)
    return full if self.S else full[0]


class StreamedNB(D.NegativeBinomialDisp):
  """NB(mean, inverse dispersion) of the output head, parameters recomputed on demand."""

  def __init__(self, src: StreamSource, name="NegativeBinomialDisp"):
    self._src, self.name = src, name

  @property
  def batch_shape(self):
    return self._src.shape(self._src.eng.cfg.n_genes)

  @property
  def loc(self):
    return self._src.gather("mean", self._src.eng.cfg.n_genes, want_mean=True)

  @property
  def disp(self):
    return self._src.gather("disp", self._src.eng.cfg.n_genes, want_disp=True)

  def mean(self):
    return self.loc

  def variance(self):
    m, d = self.loc, self.disp
    return m + m * m / d

  def mean_over_samples(self) -> torch.Tensor:
    """[N, G]: mean over the Monte-Carlo samples of the NB mean, accumulated inside the fused epilogue."""
    G = self._src.eng.cfg.n_genes
    if not self._src.fused:        # un-fused cross-check path (gemm_mode 0): no epilogue to accumulate in
      m = self.loc
      return m.mean(dim=0) if m.dim() == 3 else m
    out = torch.empty((self._src.N, G), dtype=torch.float32, device=self._src.eng.device)
    for k, sl in self._src.chunks():
      out[sl] = self._src.run(k, sl, want_mean_avg=True, want_latent=False)["mean_avg"]
    return out

  def cell_log_prob(self, x) -> torch.Tensor:
    """[S, N] (or [N]): log-likelihood of the count rows `x` [N, G] summed over genes, zero inflation stripped."""
    return _cell_log_prob(self._src, x, strip_zi=True)

  def materialize(self) -> D.NegativeBinomialDisp:
    return D.NegativeBinomialDisp(self.loc, self.disp, self.name)

  def sample(self, sample_shape=()):
    return self.materialize().sample(sample_shape)

  def log_prob(self, x):
    return self.materialize().log_prob(x)


class StreamedZeroInflated(D.ZeroInflated):
  def __init__(self, src: StreamSource, name="ZeroInflated"):
    self._src, self.name = src, name
    self.count_distribution = StreamedNB(src)

  @property
  def logits(self):
    return self._src.gather("pi_logit", self._src.eng.cfg.n_genes, want_pi=True)

  def cell_log_prob(self, x) -> torch.Tensor:
    return _cell_log_prob(self._src, x, strip_zi=False)

  def materialize(self) -> D.ZeroInflated:
    return D.ZeroInflated(self.count_distribution.materialize(), self.logits, self.name)

  def mean(self):
    return self.materialize().mean()

  def variance(self):
    return self.materialize().variance()

  def sample(self, sample_shape=()):
    return self.materialize().sample(sample_shape)

  def log_prob(self, x):
    return self.materialize().log_prob(x)


def _cell_log_prob(src: StreamSource, x, strip_zi: bool) -> torch.Tensor:
  eng = src.eng
  x = eng._dev(x)
  if strip_zi and not src.fused and eng.cfg.n_out_heads == 3:      # un-fused cross-check path: dense parameters, eager log-prob
    nb = D.NegativeBinomialDisp(src.gather("mean", eng.cfg.n_genes, want_mean=True), src.gather("disp", eng.cfg.n_genes, want_disp=True))
    return nb.log_prob(x).sum(-1)
  if tuple(x.shape) != (src.N, eng.cfg.n_genes):
    raise ValueError(f"log_prob expects counts of shape {(src.N, eng.cfg.n_genes)}, got {tuple(x.shape)}")
  out = torch.empty((src.samples, src.N), dtype=torch.float32, device=eng.device)
  for k, sl in src.chunks():
    t = src.run(k, sl, x_eval=x, strip_zi=strip_zi, want_latent=False)["terms"]
    out[:, sl] = t[1].reshape(src.samples, sl.stop - sl.start)
  return out if src.S else out[0]


class StreamedIndependent(D.Independent):
  """``Independent(base, 1)`` over the genes: ``log_prob`` comes per cell from the fused epilogue."""

  def __init__(self, base, name=None):
    self.distribution, self.nd = base, 1
    self.name = name or base.name

  def log_prob(self, x):
    return self.distribution.cell_log_prob(x)

  def mean_over_samples(self) -> torch.Tensor:
    base = self.distribution
    nb = base.count_distribution if isinstance(base, D.ZeroInflated) else base
    return nb.mean_over_samples()

  def materialize(self) -> D.Independent:
    return D.Independent(self.distribution.materialize(), 1, name=self.name)


def imputed_distribution(pX):
  """The reference's "imputed" output: the count distribution with its zero inflation stripped, re-wrapped in
  ``Independent`` (sisua/analysis/posterior.py:210-220)."""
  base = pX.distribution if isinstance(pX, D.Independent) else pX
  if isinstance(base, D.ZeroInflated):
    base = base.count_distribution
  if isinstance(pX, StreamedIndependent):
    return StreamedIndependent(base, name=pX.name)
  return D.Independent(base, 1, name=getattr(pX, "name", None)) if isinstance(pX, D.Independent) else base
