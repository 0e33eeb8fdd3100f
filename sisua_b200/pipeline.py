"""Host-buffer entry point of the train step: the caller hands pinned host minibatches (what a
`tf.data`-style input pipeline yields, sisua/data/_single_cell_base.py:593-601); copies are
double-buffered on a side stream so the H2D transfer of step i+1 overlaps the kernels of step i."""
from __future__ import annotations

from typing import Callable, List, Optional

import torch

from .engine import Engine


class HostTrainPipeline:
  def __init__(self, eng: Engine, batch: int, depth: int = 2):
    self.eng = eng
    dev = eng.device
    cfg = eng.cfg
    self.copy_stream = torch.cuda.Stream(device=dev)
    self.depth = depth
    self.x = [torch.empty((batch, cfg.n_genes), device=dev) for _ in range(depth)]
    self.eps = [torch.empty((batch, cfg.n_latent), device=dev) for _ in range(depth)]
    self.ready = [torch.cuda.Event() for _ in range(depth)]
    self.consumed = [torch.cuda.Event() for _ in range(depth)]
    self.terms = torch.empty((5, batch), device=dev)
    self.loss = torch.empty((1,), device=dev)
    self.host_loss = [torch.empty((1,), dtype=torch.float32).pin_memory() for _ in range(depth)]
    self.i = 0
    for e in self.consumed:
      e.record()

  def step(self, x_host: torch.Tensor, eps_host: torch.Tensor, step: int, lr: float = 1e-3, clipnorm: float = 100.0,
           world: int = 1, allreduce: Optional[Callable] = None):
    """Enqueue one train step from pinned host buffers; returns the pinned host tensor that will hold
    the loss once the stream has drained (read it after `flush`)."""
    eng = self.eng
    s = self.i % self.depth
    self.i += 1
    main = torch.cuda.current_stream(eng.device)
    with torch.cuda.stream(self.copy_stream):
      self.copy_stream.wait_event(self.consumed[s])
      self.x[s].copy_(x_host, non_blocking=True)
      self.eps[s].copy_(eps_host, non_blocking=True)
      self.ready[s].record(self.copy_stream)
    main.wait_event(self.ready[s])
    eng.train_step(self.x[s], eps_z=self.eps[s], terms=self.terms, loss=self.loss, seed=0, step=step)
    self.consumed[s].record(main)
    if allreduce is not None:
      allreduce(eng.grads)
    eng.adam_step(lr=lr, clipnorm=clipnorm, grad_scale=1.0 / world, t=step)
    out = self.host_loss[s]
    out.copy_(self.loss, non_blocking=True)
    return out

  def flush(self, losses: List[torch.Tensor]):
    torch.cuda.current_stream(self.eng.device).synchronize()
    return [float(l.item()) for l in losses[-self.depth:]]
