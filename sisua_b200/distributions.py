"""Light distribution objects returned by ``SingleCellModel.predict`` / ``__call__``.

They mirror the slice of the TFP / odin-ai distribution surface the reference's callers touch
(sisua/analysis/posterior.py:210-220,919-938, tests/test_singlecell_models.py:116-188):
``mean() variance() stddev() sample(n) log_prob(x) batch_shape event_shape name`` and, for
zero-inflated outputs, ``.distribution.count_distribution``.  They only hold the parameter tensors
the CUDA step produced (torch tensors, on the GPU or pinned host); nothing here is on the hot path."""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch
import torch.nn.functional as F

EPS = 1e-8


class Distribution:
  name: str = "Distribution"

  @property
  def batch_shape(self) -> Tuple[int, ...]:
    raise NotImplementedError

  @property
  def event_shape(self) -> Tuple[int, ...]:
    return ()

  def mean(self):
    raise NotImplementedError

  def variance(self):
    raise NotImplementedError

  def stddev(self):
    return torch.sqrt(self.variance())

  def sample(self, sample_shape=()):
    raise NotImplementedError

  def log_prob(self, x):
    raise NotImplementedError

  @staticmethod
  def _shape(sample_shape):
    if isinstance(sample_shape, int):
      return (sample_shape,)
    return tuple(sample_shape)

  def __repr__(self):
    return (f"<{type(self).__name__} '{self.name}' batch_shape={tuple(self.batch_shape)} "
            f"event_shape={tuple(self.event_shape)}>")


class Normal(Distribution):
  def __init__(self, loc, scale, name="Normal"):
    self.loc, self.scale, self.name = loc, scale, name

  @property
  def batch_shape(self):
    return tuple(self.loc.shape)

  def mean(self):
    return self.loc

  def variance(self):
    return self.scale * self.scale

  def sample(self, sample_shape=()):
    s = self._shape(sample_shape)
    return self.loc + self.scale * torch.randn(s + tuple(self.loc.shape), device=self.loc.device)

  def log_prob(self, x):
    x = torch.as_tensor(x, device=self.loc.device, dtype=self.loc.dtype)
    z = (x - self.loc) / self.scale
    return -0.5 * z * z - torch.log(self.scale) - 0.5 * math.log(2 * math.pi)


class NegativeBinomialDisp(Distribution):
  """NB parameterised by mean and inverse dispersion (odin-ai NegativeBinomialDisp)."""

  def __init__(self, loc, disp, name="NegativeBinomialDisp"):
    self.loc, self.disp, self.name = loc, disp, name

  @property
  def batch_shape(self):
    return tuple(self.loc.shape)

  def mean(self):
    return self.loc

  def variance(self):
    return self.loc + self.loc * self.loc / self.disp

  def sample(self, sample_shape=()):
    s = self._shape(sample_shape)
    shape = s + tuple(self.loc.shape)
    rate = torch.distributions.Gamma(self.disp.expand(shape), (self.disp / (self.loc + EPS)).expand(shape)).sample()
    return torch.poisson(rate)

  def log_prob(self, x):
    x = torch.as_tensor(x, device=self.loc.device, dtype=self.loc.dtype)
    mu, th = self.loc, self.disp
    ltm = torch.log(th + mu + EPS)
    return (th * (torch.log(th + EPS) - ltm) + x * (torch.log(mu + EPS) - ltm) + torch.lgamma(x + th) -
            torch.lgamma(th) - torch.lgamma(x + 1.0))


class NegativeBinomial(Distribution):
  """TFP NegativeBinomial(total_count, logits) — the 'nb' protein head (configs/base.yaml:38-40)."""

  def __init__(self, total_count, logits, name="NegativeBinomial"):
    self.total_count, self.logits, self.name = total_count, logits, name

  @property
  def batch_shape(self):
    return tuple(self.total_count.shape)

  def mean(self):
    return self.total_count * torch.exp(self.logits)

  def variance(self):
    return self.mean() / torch.sigmoid(-self.logits)

  def sample(self, sample_shape=()):
    s = self._shape(sample_shape)
    shape = s + tuple(self.total_count.shape)
    rate = torch.distributions.Gamma(self.total_count.expand(shape), torch.exp(-self.logits).expand(shape)).sample()
    return torch.poisson(rate)

  def log_prob(self, x):
    x = torch.as_tensor(x, device=self.logits.device, dtype=self.logits.dtype)
    r = self.total_count
    return (torch.lgamma(r + x) - torch.lgamma(r) - torch.lgamma(x + 1.0) + r * F.logsigmoid(-self.logits) +
            x * F.logsigmoid(self.logits))


class ZeroInflated(Distribution):
  """Mixture of a point mass at zero (probability sigmoid(logits)) and ``count_distribution``."""

  def __init__(self, count_distribution: Distribution, logits, name="ZeroInflated"):
    self.count_distribution, self.logits, self.name = count_distribution, logits, name

  @property
  def batch_shape(self):
    return self.count_distribution.batch_shape

  @property
  def probs(self):
    return torch.sigmoid(self.logits)

  def mean(self):
    return torch.sigmoid(-self.logits) * self.count_distribution.mean()

  def variance(self):
    q = torch.sigmoid(-self.logits)
    m, v = self.count_distribution.mean(), self.count_distribution.variance()
    return q * (v + m * m) - (q * m) ** 2

  def sample(self, sample_shape=()):
    c = self.count_distribution.sample(sample_shape)
    keep = torch.rand_like(c) >= torch.sigmoid(self.logits)
    return c * keep

  def log_prob(self, x):
    x = torch.as_tensor(x, device=self.logits.device, dtype=self.logits.dtype)
    base = self.count_distribution.log_prob(x)
    pi = self.logits
    zero = F.softplus(base - pi) - F.softplus(-pi)     # log(sig(pi) + sig(-pi) * p0)
    nonzero = -F.softplus(pi) + base
    return torch.where(x < EPS, zero, nonzero)


class Independent(Distribution):
  """Sums the last ``reinterpreted_batch_ndims`` batch axes into the event."""

  def __init__(self, distribution: Distribution, reinterpreted_batch_ndims: int = 1, name=None):
    self.distribution = distribution
    self.nd = int(reinterpreted_batch_ndims)
    self.name = name or distribution.name

  @property
  def batch_shape(self):
    b = self.distribution.batch_shape
    return tuple(b[:len(b) - self.nd])

  @property
  def event_shape(self):
    b = self.distribution.batch_shape
    return tuple(b[len(b) - self.nd:])

  def mean(self):
    return self.distribution.mean()

  def variance(self):
    return self.distribution.variance()

  def sample(self, sample_shape=()):
    return self.distribution.sample(sample_shape)

  def log_prob(self, x):
    lp = self.distribution.log_prob(x)
    return lp.sum(dim=tuple(range(-self.nd, 0)))

  @property
  def is_zero_inflated(self):
    return isinstance(self.distribution, ZeroInflated)


class MultivariateNormalDiag(Independent):
  def __init__(self, loc, scale_diag, name="MultivariateNormalDiag"):
    super().__init__(Normal(loc, scale_diag, name), 1, name)
    self.loc, self.scale_diag = loc, scale_diag

  def kl_standard_normal(self):
    s, m = self.scale_diag, self.loc
    return 0.5 * torch.sum(s * s + m * m - 1.0 - 2.0 * torch.log(s), dim=-1)


class VectorDeterministic(Distribution):
  """Deterministic latent of the Deep Count Autoencoder (sisua/models/dca.py:16-28)."""

  def __init__(self, loc, name="VectorDeterministic"):
    self.loc, self.name = loc, name

  @property
  def batch_shape(self):
    return tuple(self.loc.shape[:-1])

  @property
  def event_shape(self):
    return tuple(self.loc.shape[-1:])

  def mean(self):
    return self.loc

  def variance(self):
    return torch.zeros_like(self.loc)

  def sample(self, sample_shape=()):
    s = self._shape(sample_shape)
    return self.loc.expand(s + tuple(self.loc.shape))

  def log_prob(self, x):
    x = torch.as_tensor(x, device=self.loc.device, dtype=self.loc.dtype)
    return torch.where((x == self.loc).all(dim=-1), 0.0, -float("inf"))


class MeanOnly(Distribution):
  """Carrier for a head of which only the mean was requested from the CUDA step."""

  def __init__(self, loc, name="MeanOnly"):
    self.loc, self.name = loc, name

  @property
  def batch_shape(self):
    return tuple(self.loc.shape)

  def mean(self):
    return self.loc

  def log_prob(self, x):
    raise NotImplementedError("only the mean of this head is materialised; per-cell log-likelihood is in elbo_terms")
