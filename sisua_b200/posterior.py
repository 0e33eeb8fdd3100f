"""Fast paths of the inference consumer: the quantities
``sisua.analysis.posterior.Posterior`` derives from ``scm.predict`` — imputed mean (NB mean without
zero inflation, averaged over MC samples: posterior.py:210-220,986-988), latent means
(:244-253), log-likelihood of original / corrupted counts (:919-938), marginal llk (:941-976)."""
from __future__ import annotations

import numpy as np
import torch

from . import distributions as D


def corrupt_binomial(X: np.ndarray, dropout_rate=0.2, retain_rate=0.2, seed=1) -> np.ndarray:
  """Binomial down-sampling of a random subset of entries (sisua/data/utils.py:168-228 semantics:
  a fraction ``dropout_rate`` of the entries is replaced by Binomial(x, retain_rate))."""
  rng = np.random.RandomState(seed)
  X = np.array(X, dtype=np.float32, copy=True)
  sel = rng.random_sample(X.shape) < dropout_rate
  X[sel] = rng.binomial(X[sel].astype(np.int64), retain_rate).astype(np.float32)
  return X


class Posterior:
  def __init__(self, scm, sco, dropout_rate=0.2, retain_rate=0.2, corrupt_distribution='binomial', batch_size=8,
               sample_shape=10, random_state=1, name=None):
    from .models import SingleCellData
    if corrupt_distribution != 'binomial':
      raise ValueError("only the 'binomial' corruption of the reference default is implemented")
    self.scm, self.name = scm, name or "posterior"
    self.sco_original = sco
    Xc = corrupt_binomial(sco.X, dropout_rate, retain_rate, random_state)
    self.sco_corrupted = SingleCellData(Xc, sco.Y, name=sco.name + "_corrupted", var_names=sco.var_names)
    self.sco_corrupted.mask = sco.mask
    self.sample_shape = sample_shape
    pX, qZ = scm.predict(self.sco_corrupted, sample_shape=sample_shape, batch_size=batch_size, verbose=False)
    self.pX = pX[0] if isinstance(pX, tuple) else pX
    self.qZ = qZ[0] if isinstance(qZ, tuple) else qZ

  @property
  def imputed(self) -> torch.Tensor:
    """Mean over MC samples of the count distribution's mean, zero inflation stripped."""
    base = self.pX.distribution
    nb = base.count_distribution if isinstance(base, D.ZeroInflated) else base
    m = nb.mean()
    return m.mean(dim=0) if m.dim() == 3 else m

  @property
  def latents(self) -> torch.Tensor:
    return self.qZ.mean()

  def cal_llk(self):
    """log mean_s p(x | z_s) per cell on the original and the corrupted counts."""
    out = {}
    S = self.pX.batch_shape[0] if len(self.pX.batch_shape) == 2 else 1
    for tag, sco in (("original", self.sco_original), ("corrupted", self.sco_corrupted)):
      x = torch.from_numpy(sco.X).to(self.pX.mean().device)
      lp = self.pX.log_prob(x)
      if lp.dim() == 2:
        lp = torch.logsumexp(lp, dim=0) - float(np.log(S))
      out[tag] = float(lp.mean())
    return out

  def cal_imputation_scores(self):
    """d(original, imputed) vs d(original, corrupted): mean absolute error on the corrupted entries."""
    X, Xc = self.sco_original.X, self.sco_corrupted.X
    sel = X != Xc
    imp = self.imputed.cpu().numpy()
    if not sel.any():
      return dict(imputed=0.0, corrupted=0.0)
    return dict(imputed=float(np.abs(imp[sel] - X[sel]).mean()), corrupted=float(np.abs(Xc[sel] - X[sel]).mean()))
