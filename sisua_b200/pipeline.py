"""Host-buffer entry point of the train step: the caller hands pinned host minibatches (what a
`tf.data`-style input pipeline yields, sisua/data/_single_cell_base.py:593-601); copies are
double-buffered on a side stream so the H2D transfer of step i+1 overlaps the kernels of step i."""
from __future__ import annotations

from typing import Callable, List, Optional

import torch

from .engine import Engine


def quantize_counts(X):
  """float32 count matrix -> pinned uint16 host tensor when every value is an integer below 65536 (done once per
  dataset, like the reference's cached tf.data pipeline); otherwise the pinned float32 tensor."""
  import numpy as np
  X = np.ascontiguousarray(X)
  if X.dtype != np.uint16 and (X.min() < 0 or X.max() >= 65536 or not np.array_equal(X, np.rint(X))):
    return torch.from_numpy(X.astype(np.float32, copy=False)).pin_memory()
  return torch.from_numpy(X.astype(np.uint16).view(np.int16)).pin_memory()


class CsrBatch:
  """Pinned host CSR form of one minibatch of integer counts (built once per dataset / epoch cache)."""

  def __init__(self, X):
    import numpy as np
    X = np.ascontiguousarray(X)
    if X.shape[1] > 65536 or X.min() < 0 or X.max() >= 65536 or not np.array_equal(X, np.rint(X)):
      raise ValueError("CsrBatch needs non-negative integer counts below 65536 and at most 65536 genes")
    r, c = np.nonzero(X)
    self.rows, self.genes = X.shape
    self.indptr = torch.from_numpy(np.concatenate([[0], np.cumsum(np.bincount(r, minlength=X.shape[0]))]).astype(np.int32)).pin_memory()
    self.cols = torch.from_numpy(c.astype(np.uint16).view(np.int16)).pin_memory()
    self.vals = torch.from_numpy(X[r, c].astype(np.uint16).view(np.int16)).pin_memory()

  @property
  def nbytes(self) -> int:
    return self.indptr.numel() * 4 + self.cols.numel() * 2 + self.vals.numel() * 2


class HostTrainPipeline:
  def __init__(self, eng: Engine, batch: int, depth: int = 2):
    self.eng = eng
    dev = eng.device
    cfg = eng.cfg
    self.copy_stream = torch.cuda.Stream(device=dev)
    self.depth = depth
    self.x = [torch.empty((batch, cfg.n_genes), device=dev) for _ in range(depth)]
    self.x16 = [torch.empty((batch, cfg.n_genes), device=dev, dtype=torch.int16) for _ in range(depth)]
    self.csr = [None] * depth   # (indptr, cols, vals) device buffers, grown on demand
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
      is_csr = isinstance(x_host, CsrBatch)
      packed = (not is_csr) and x_host.dtype in (torch.int16, torch.uint16)   # integer counts shipped as 16-bit (see quantize_counts)
      if is_csr:
        nnz = x_host.cols.numel()
        if self.csr[s] is None or self.csr[s][1].numel() < nnz:
          cap = int(nnz * 1.25) + 1024
          self.csr[s] = (torch.empty(self.x[s].shape[0] + 1, device=eng.device, dtype=torch.int32),
                         torch.empty(cap, device=eng.device, dtype=torch.int16), torch.empty(cap, device=eng.device, dtype=torch.int16))
        ip, cc, vv = self.csr[s]
        ip.copy_(x_host.indptr, non_blocking=True)
        cc[:nnz].copy_(x_host.cols, non_blocking=True)
        vv[:nnz].copy_(x_host.vals, non_blocking=True)
      elif packed:
        self.x16[s].copy_(x_host.view(torch.int16), non_blocking=True)
      else:
        self.x[s].copy_(x_host, non_blocking=True)
      self.eps[s].copy_(eps_host, non_blocking=True)
      self.ready[s].record(self.copy_stream)
    main.wait_event(self.ready[s])
    if is_csr:
      eng.unpack_counts_csr(self.csr[s][0], self.csr[s][1], self.csr[s][2], self.x[s])
    elif packed:
      eng.unpack_counts_u16(self.x16[s], self.x[s])
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


class GraphedTrainStep:
  """train_step + adam_step captured once into a CUDA graph and replayed per minibatch: at the reference's
  minibatch sizes (64 / 128, configs/base.yaml:23) the step is launch-bound, so replaying ~20 kernels as one graph
  launch is what sets the rate.  Inputs are copied into static device buffers; dropout masks and the Adam bias
  correction follow the device-side step counter, so every replay is a fresh step."""

  def __init__(self, eng: Engine, batch: int, lr: float = 1e-3, clipnorm: float = 100.0, seed: int = 0,
               with_y: bool = False, with_library: bool = False):
    cfg = eng.cfg
    dev = eng.device
    self.eng, self.batch = eng, batch
    self.x = torch.zeros((batch, cfg.n_genes), device=dev)
    self.eps_z = torch.zeros((batch, cfg.n_latent), device=dev) if cfg.model_kind != 2 else None
    self.eps_l = torch.zeros((batch,), device=dev) if cfg.model_kind == 1 else None
    self.library = torch.ones((batch, 2), device=dev) if cfg.model_kind == 1 else None
    self.y = torch.zeros((batch, cfg.n_proteins), device=dev) if cfg.n_proteins > 0 else None
    self.mask = torch.zeros((batch,), device=dev, dtype=torch.uint8) if cfg.n_proteins > 0 else None
    self.terms = torch.empty((5, batch), device=dev)
    self.loss = torch.empty((1,), device=dev)
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):            # warm-up outside capture, then restore the optimiser state it touched
      snap = [t.clone() for t in (eng.params, eng.adam_m, eng.adam_v, eng.bn_moving)]
      step0 = eng.step_count
      self._run(lr, clipnorm, seed)
      side.synchronize()
      for t, s_ in zip((eng.params, eng.adam_m, eng.adam_v, eng.bn_moving), snap):
        t.copy_(s_)
      eng.reset_step_counter(step0)
    torch.cuda.current_stream(dev).wait_stream(side)
    self.graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(self.graph):
      self._run(lr, clipnorm, seed)
    eng.reset_step_counter(step0)            # capture does not execute, but keep host / device counters aligned

  def _run(self, lr, clipnorm, seed):
    self.eng.train_step(self.x, y=self.y, library=self.library, mask=self.mask, eps_z=self.eps_z, eps_l=self.eps_l,
                        terms=self.terms, loss=self.loss, seed=seed, step=-1)
    self.eng.adam_step(lr=lr, clipnorm=clipnorm, grad_scale=1.0, t=0)

  def step(self, x, eps_z=None, eps_l=None, library=None, y=None, mask=None):
    """Copies the minibatch into the static buffers and replays the graph; results land in self.terms / self.loss."""
    self.x.copy_(x, non_blocking=True)
    for dst, src in ((self.eps_z, eps_z), (self.eps_l, eps_l), (self.library, library), (self.y, y), (self.mask, mask)):
      if dst is not None and src is not None:
        dst.copy_(src, non_blocking=True)
    self.graph.replay()
    self.eng.step_count += 1
    return self.terms, self.loss
