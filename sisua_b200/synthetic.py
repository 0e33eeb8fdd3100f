"""Deterministic synthetic single-cell count generators.

Shapes / statistics are calibrated to the datasets the reference documents in
description/dataset.html (cortex :32, pbmc8k_ly :188-189, pbmc5000 :158); the dense
``stress`` generator is the reference's own scalability input
(tests/test_scalability.py:21-28,64-65: randint(0,100) counts, randint(0,10) proteins,
SEED 87654321)."""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np

DATA_SEED = 87654321     # tests/test_scalability.py:21
WEIGHT_SEED = 8          # sisua/train.py:24-25

PRESETS = {
    # name: (n_cells, n_genes, n_proteins, mean_log_library, sd_log_library)
    "cortex": (3005, 558, 0, 8.46, 0.66),
    "pbmc8k": (8381, 2000, 0, 6.42, 0.28),
    "pbmc8k_ly": (8381, 500, 10, 6.42, 0.28),
    "dca100k": (100_000, 5000, 0, 6.0, 0.5),
    "scale1m": (1_000_000, 2000, 0, 6.42, 0.28),
    "scale1m_20k": (1_000_000, 20000, 0, 6.42, 0.28),
}


def realistic_counts(n_cells: int, n_genes: int, n_proteins: int = 0, mean_log: float = 6.42,
                     sd_log: float = 0.28, seed: int = DATA_SEED, dropout: float = 0.1,
                     chunk: int = 65536) -> Dict[str, np.ndarray]:
  """x_cg ~ NB(mean = l_c m_g, theta_g) then zeroed w.p. ``dropout``; float32 counts."""
  rng = np.random.default_rng(seed)
  log_m = rng.normal(0.0, 1.5, size=n_genes)
  m = np.exp(log_m - log_m.max())
  m /= m.sum()
  theta = rng.gamma(2.0, 1.0, size=n_genes) + 1e-3
  X = np.empty((n_cells, n_genes), dtype=np.float32)
  lib = np.empty(n_cells, dtype=np.float64)
  for s in range(0, n_cells, chunk):
    e = min(n_cells, s + chunk)
    l = rng.lognormal(mean_log, sd_log, size=e - s)
    lib[s:e] = l
    mean = l[:, None] * m[None, :]
    lam = rng.gamma(theta[None, :], mean / theta[None, :])
    c = rng.poisson(lam).astype(np.float32)
    c[rng.random(c.shape) < dropout] = 0.0
    X[s:e] = c
  out = {"x": X}
  if n_proteins > 0:
    w = rng.lognormal(0.0, 1.0, size=n_proteins)
    mean = 50.0 * w[None, :] * (lib / lib.mean())[:, None]
    lam = rng.gamma(5.0, mean / 5.0)
    out["y"] = np.log1p(rng.poisson(lam)).astype(np.float32)   # continuous, ~[0, 9]
  return out


def stress_counts(n_cells: int, n_genes: int = 500, n_proteins: int = 10,
                  seed: int = DATA_SEED) -> Dict[str, np.ndarray]:
  rng = np.random.RandomState(seed % (2 ** 32))
  out = {"x": rng.randint(0, 100, size=(n_cells, n_genes)).astype(np.float32)}
  if n_proteins > 0:
    out["y"] = rng.randint(0, 10, size=(n_cells, n_proteins)).astype(np.float32)
  return out


def preset(name: str, n_cells: Optional[int] = None, seed: int = DATA_SEED) -> Dict[str, np.ndarray]:
  n, g, p, ml, sd = PRESETS[name]
  return realistic_counts(n_cells or n, g, p, ml, sd, seed=seed)


def library_stats(X: np.ndarray) -> np.ndarray:
  """get_library_size (sisua/data/utils.py:231-263): [N,2] = (mean, var) of log(sum_g x + 1e-8),
  dataset-level constants broadcast to each cell (the `library` entry of the minibatch dict,
  sisua/data/_single_cell_base.py:568-570)."""
  total = X.sum(axis=1, dtype=np.float64)
  lc = np.log(total + 1e-8)
  out = np.empty((X.shape[0], 2), dtype=np.float32)
  out[:, 0] = lc.mean()
  out[:, 1] = lc.var()
  return out


def label_mask(n_cells: int, labels_percent: float = 0.1, seed: int = 1) -> np.ndarray:
  """Frozen per-cell semi-supervision mask (sisua/data/_single_cell_base.py:580-591: drawn once,
  generator seed 1, then cached for all epochs)."""
  rng = np.random.default_rng(seed)
  return (rng.random(n_cells) < labels_percent).astype(np.uint8)
