"""Fast paths of the inference consumer: what ``sisua.analysis.posterior.Posterior`` derives
from ``scm.predict`` -- imputed / reconstructed distributions (posterior.py:210-220), latent means (:244-253),
log-likelihood of original / corrupted counts under both (:919-938), marginal log-likelihood (:941-976), imputation
scores (:979-993 with sisua/analysis/imputation_benchmarks.py:102-130).

Everything gene-sized runs inside the fused kernels through the streamed distributions of ``sisua_b200.streamed``
(``sisua_infer_ex`` / ``sisua_marginal_llk``); what is left on the host are reductions over ``[S, N]`` / ``[N]`` vectors
and the medians of the imputation scores."""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch

from . import distributions as D
from . import streamed as ST


def apply_artificial_corruption(X: np.ndarray, dropout: float = 0.0, distribution: str = "binomial", retain_rate: float = 0.2,
                                seed: int = 8) -> np.ndarray:
  """Host restatement of sisua/data/utils.py:168-228 (scVI's corruption recipe), same RandomState call sequence: a fraction
  ``dropout`` of the NON-ZERO entries is drawn without replacement; 'binomial' replaces each by Binomial(x, retain_rate),
  'uniform' multiplies it by Bernoulli(retain_rate).  Returns a corrupted copy."""
  distribution = str(distribution).lower()
  dropout = float(dropout)
  assert 0 <= dropout < 1, f"dropout value must be >= 0 and < 1, given: {dropout}"
  X = np.array(X, dtype=np.float32, copy=True)
  if not (0. < dropout < 1. or 0. < retain_rate < 1.):
    return X
  rand = np.random.RandomState(seed=seed)
  i, j = np.nonzero(X)
  ix = rand.choice(range(len(i)), size=int(np.floor(dropout * len(i))), replace=False)
  i, j = i[ix], j[ix]
  if distribution == "uniform":
    corrupted = np.multiply(X[i, j], rand.binomial(n=np.ones(len(ix), dtype=np.int32), p=retain_rate))
  elif distribution == "binomial":
    corrupted = rand.binomial(n=(X[i, j]).astype(np.int32), p=retain_rate)
  else:
    raise ValueError(f"Only support 2 corruption distribution: 'uniform' and 'binomial', but given: '{distribution}'")
  X[i, j] = corrupted
  return X


def corrupt_on_device(engine, X, dropout: float = 0.0, distribution: str = "binomial", retain_rate: float = 0.2, seed: int = 8,
                      rows_per_call: int = 1 << 18) -> np.ndarray:
  """``apply_artificial_corruption`` on the GPU (``sisua_corrupt_counts``): the matrix is streamed through the device in row
  blocks; row indices of the Philox counters are global, so the result does not depend on the blocking."""
  X = np.asarray(X, dtype=np.float32)
  assert 0 <= dropout < 1, f"dropout value must be >= 0 and < 1, given: {dropout}"
  if not (0. < dropout < 1. or 0. < retain_rate < 1.):
    return X.copy()
  if X.shape[0] <= rows_per_call:
    return engine.corrupt_counts(torch.from_numpy(X), dropout, retain_rate, distribution, seed).cpu().numpy()
  raise NotImplementedError("corrupt_on_device: matrices above 2^18 rows -- corrupt the shard that is resident with Engine.corrupt_counts")


corrupt_binomial = lambda X, dropout_rate=0.2, retain_rate=0.2, seed=1: apply_artificial_corruption(   # round-1 name
    X, dropout=dropout_rate, distribution="binomial", retain_rate=retain_rate, seed=seed)


def imputation_score(original: np.ndarray, imputed: np.ndarray) -> float:
  """Median of all distances (imputation_benchmarks.py:102-107)."""
  assert original.shape == imputed.shape
  return float(np.median(np.abs(original - imputed)))


def _per_cell_medians(original, corrupted, imputed) -> np.ndarray:
  changed = original.sum(axis=1) != corrupted.sum(axis=1)
  if not changed.any():
    return np.zeros(0)
  return np.median(np.abs(original[changed] - imputed[changed]), axis=1)


def imputation_mean_score(original, corrupted, imputed) -> float:
  """Mean of the per-cell medians over the cells the corruption touched (imputation_benchmarks.py:110-118)."""
  m = _per_cell_medians(original, corrupted, imputed)
  return float(m.mean()) if m.size else 0.0


def imputation_std_score(original, corrupted, imputed) -> float:
  m = _per_cell_medians(original, corrupted, imputed)
  return float(m.std()) if m.size else 0.0


class Posterior:
  r""" Posterior of a fitted ``SingleCellModel`` on a test set: the test counts are corrupted, the model runs on the
  corrupted counts with ``sample_shape`` Monte-Carlo samples, and the scores compare against the original counts
  (sisua/analysis/posterior.py:108-255). """

  def __init__(self, scm, sco, dropout_rate=0.2, retain_rate=0.2, corrupt_distribution='binomial', batch_size=8,
               sample_shape=10, random_state=1, name=None, verbose=False, corrupt_on='host'):
    from .models import SingleCellData
    if not scm.is_fitted:
      raise RuntimeError("fit() must be called before creating Posterior.")
    self.scm, self.name, self.verbose = scm, name or "posterior", verbose
    self.sco_original = sco
    if corrupt_on == 'host':      # the reference's NumPy routine, same RandomState sequence
      Xc = apply_artificial_corruption(sco.X, dropout=dropout_rate, distribution=corrupt_distribution, retain_rate=retain_rate,
                                       seed=random_state)
    elif corrupt_on == 'device':  # Philox on the GPU (sisua_corrupt_counts): per-entry selection, see include/sisua_b200.h
      Xc = corrupt_on_device(scm.engine, sco.X, dropout=dropout_rate, distribution=corrupt_distribution, retain_rate=retain_rate,
                             seed=random_state)
    else:
      raise ValueError(f"corrupt_on must be 'host' or 'device', given: {corrupt_on!r}")
    self.sco_corrupted = SingleCellData(Xc, sco.Y, name=sco.name + "_corrupted", var_names=sco.var_names)
    self.sco_corrupted.mask = sco.mask
    self.sample_shape = sample_shape
    self.omic = scm.output_layers[0].name
    pX, qZ = scm.predict(self.sco_corrupted, sample_shape=sample_shape, batch_size=batch_size, verbose=False)
    pX0 = pX[0] if isinstance(pX, tuple) else pX
    self.qZ = qZ[0] if isinstance(qZ, tuple) else qZ
    # (name, 'reconstructed') = the model's output distribution; (name, 'imputed') = its count distribution without
    # the zero inflation (posterior.py:210-220)
    self.omics_data = {(self.omic, "reconstructed"): pX0, (self.omic, "imputed"): ST.imputed_distribution(pX0),
                       (self.omic, "original"): sco.X, (self.omic, "corrupted"): Xc}
    self.pX = pX0
    self._cache: Dict[str, object] = {}

  # -------------------------------------------------------------------------------------------
  @property
  def imputed(self) -> torch.Tensor:
    """[N, G] mean over the Monte-Carlo samples of the count distribution's mean (zero inflation stripped), accumulated
    in the fused epilogue."""
    if "imputed" not in self._cache:
      d = self.omics_data[(self.omic, "imputed")]
      if isinstance(d, ST.StreamedIndependent):
        self._cache["imputed"] = d.mean_over_samples()
      else:
        m = d.mean()
        self._cache["imputed"] = m.mean(dim=0) if m.dim() == 3 else m
    return self._cache["imputed"]

  @property
  def latents(self) -> torch.Tensor:
    return self.qZ.mean()

  def cal_llk(self, omic=None) -> Dict[str, float]:
    r""" Log-likelihood of the original / corrupted counts under the imputed / reconstructed distributions: per cell
    ``logsumexp_s llk - log S``, then the mean over cells -- the four keys of posterior.py:919-938. """
    name = self.omic
    rec, imp = self.omics_data[(name, "reconstructed")], self.omics_data[(name, "imputed")]
    x_org, x_cor = self.sco_original.X, self.sco_corrupted.X

    def reduce(llk: torch.Tensor) -> float:
      if llk.dim() == 2:
        llk = torch.logsumexp(llk, dim=0) - float(np.log(llk.shape[0]))
      return float(llk.mean())

    return {f"llk_{name}_imp_org": reduce(imp.log_prob(x_org)), f"llk_{name}_imp_cor": reduce(imp.log_prob(x_cor)),
            f"llk_{name}_rec_cor": reduce(rec.log_prob(x_cor)), f"llk_{name}_rec_org": reduce(rec.log_prob(x_org))}

  def cal_marginal_llk(self, sample_shape=100, batch_size: Optional[int] = None) -> Dict[str, float]:
    r""" Marginal log-likelihood (importance-weighted, ``sample_shape`` samples) and the reconstruction log-likelihood
    of the ORIGINAL test set: ``{"<output>_llk": .., "marginal_llk": ..}`` (posterior.py:941-976; the reference feeds
    minibatches of 2 cells with every label observed, ``labels_percent=1.0``). """
    sco, scm = self.sco_original, self.scm
    eng = scm.engine
    bs = batch_size or max(1, eng.cfg.max_batch // int(sample_shape))
    mllk, llk = [], {}
    for s in range(0, len(sco), bs):
      sl = slice(s, min(len(sco), s + bs))
      inputs = sco.X[sl] if not scm.labels else (sco.X[sl], sco.Y[sl])
      mask = np.ones(sl.stop - sl.start, dtype=np.uint8) if scm.labels else None
      library = sco.library[sl] if eng.cfg.model_kind == 1 else None
      m, d = scm.marginal_log_prob(inputs, library=library, mask=mask, sample_shape=sample_shape)
      mllk.append(m)
      for k, v in d.items():
        llk.setdefault(k, []).append(v)
    out = {f"{k}_llk": float(torch.cat(v).mean()) for k, v in llk.items()}
    out["marginal_llk"] = float(torch.cat(mllk).mean())
    return out

  def cal_imputation_scores(self) -> Dict[str, float]:
    r""" Distance between the original and the imputed counts, smaller is better (posterior.py:979-993). """
    X_org, X_crr = self.sco_original.X, self.sco_corrupted.X
    imputed = self.imputed.cpu().numpy()
    return {"imputation_med": imputation_score(X_org, imputed),
            "imputation_mean": imputation_mean_score(X_org, X_crr, imputed),
            "imputation_std": imputation_std_score(X_org, X_crr, imputed)}
