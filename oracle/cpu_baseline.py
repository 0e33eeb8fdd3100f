"""Timed CPU baseline: the oracle's train step (torch fp32, autograd, TF-formulation Adam) on the
host cores.  Stands in for the reference's TensorFlow-CPU path, which cannot be installed here
(odin-ai / TensorFlow absent, no network: DESIGN.md).  TEST / BENCH INFRASTRUCTURE ONLY."""
from __future__ import annotations

import os
import time

import numpy as np
import torch

from oracle import step_oracle as O


def host_threads() -> int:
  try:
    return len(os.sched_getaffinity(0))
  except AttributeError:
    return os.cpu_count() or 1


def time_train_steps(cfg, params_np, moving_np, batches, steps: int, warmup: int = 1, lr: float = 1e-3):
  """batches: list of oracle-style batch dicts (cycled). Returns (seconds_per_step list, threads)."""
  threads = host_threads()
  torch.set_num_threads(threads)
  P = {k: torch.tensor(np.array(v), dtype=torch.float32) for k, v in params_np.items()}
  mov = {k: torch.tensor(np.array(v), dtype=torch.float32) for k, v in moving_np.items()}
  m = {k: torch.zeros_like(v) for k, v in P.items()}
  v = {k: torch.zeros_like(p) for k, p in P.items()}
  times = []
  for t in range(1, warmup + steps + 1):
    b = batches[(t - 1) % len(batches)]
    t0 = time.perf_counter()
    drop = None
    if cfg.input_dropout > 0 or cfg.enc_dropout > 0 or cfg.dec_dropout > 0:   # fresh Bernoulli masks, drawn inside the timed step
      B = b["x"].shape[0]
      drop = {}
      if cfg.input_dropout > 0:
        drop["input"] = (torch.rand((B, cfg.n_genes)) >= cfg.input_dropout).float()
      for i in range(cfg.n_enc_layers):
        if cfg.enc_dropout > 0:
          drop[f"enc.{i}"] = (torch.rand((B, cfg.n_hidden)) >= cfg.enc_dropout).float()
      for i in range(cfg.n_dec_layers):
        if cfg.dec_dropout > 0:
          drop[f"dec.{i}"] = (torch.rand((B, cfg.n_hidden)) >= cfg.dec_dropout).float()
    O.train_step(cfg, P, mov, m, v, t, b, lr=lr, clipnorm=100.0, drop=drop)
    dt = time.perf_counter() - t0
    if t > warmup:
      times.append(dt)
  return times, threads
