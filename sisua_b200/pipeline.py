"""Host-side helpers around the train step: pinned host minibatch formats (what a `tf.data`-style input pipeline
yields, sisua/data/_single_cell_base.py:593-601), the host-buffer training loop, and the CUDA-graph replay of the
launch-bound small-batch step."""
from __future__ import annotations

from typing import Callable, List, Optional

import torch

from .engine import Engine


kRing = 8      # pinned row-pointer buffers a HostDataset cycles through; HostTrainPipeline keeps at most `depth` (< kRing) steps in flight


def _pin(t: torch.Tensor) -> torch.Tensor:
  """Page-locked when a CUDA device is present (asynchronous H2D); plain host memory otherwise, so the format helpers
  also work on a machine without a GPU (tests, data preparation)."""
  return t.pin_memory() if torch.cuda.is_available() else t


def quantize_counts(X):
  """float32 count matrix -> pinned uint16 host tensor when every value is an integer below 65536 (done once per
  dataset, like the reference's cached tf.data pipeline); otherwise the pinned float32 tensor."""
  import numpy as np
  X = np.ascontiguousarray(X)
  if X.dtype != np.uint16 and (X.min() < 0 or X.max() >= 65536 or not np.array_equal(X, np.rint(X))):
    return _pin(torch.from_numpy(X.astype(np.float32, copy=False)))
  return _pin(torch.from_numpy(X.astype(np.uint16).view(np.int16)))


class CsrBatch:
  """Pinned host CSR form of one minibatch of integer counts (built once per dataset / epoch cache)."""

  def __init__(self, X):
    import numpy as np
    X = np.ascontiguousarray(X)
    if X.shape[1] > 65536 or X.min() < 0 or X.max() >= 65536 or not np.array_equal(X, np.rint(X)):
      raise ValueError("CsrBatch needs non-negative integer counts below 65536 and at most 65536 genes")
    r, c = np.nonzero(X)
    self.rows, self.genes = X.shape
    self.indptr = _pin(torch.from_numpy(np.concatenate([[0], np.cumsum(np.bincount(r, minlength=X.shape[0]))]).astype(np.int32)))
    self.cols = _pin(torch.from_numpy(c.astype(np.uint16).view(np.int16)))
    self.vals = _pin(torch.from_numpy(X[r, c].astype(np.uint16).view(np.int16)))

  @classmethod
  def view(cls, rows: int, genes: int, indptr: torch.Tensor, cols: torch.Tensor, vals: torch.Tensor) -> "CsrBatch":
    """A minibatch that is a row range of a larger pinned CSR matrix: `indptr` [rows + 1] int32 rebased to 0, `cols` /
    `vals` the matching 16-bit slices (no copies)."""
    b = cls.__new__(cls)
    b.rows, b.genes, b.indptr, b.cols, b.vals = rows, genes, indptr, cols, vals
    return b

  @property
  def nbytes(self) -> int:
    return self.indptr.numel() * 4 + self.cols.numel() * 2 + self.vals.numel() * 2


def encode_csr8(X):
  """Packed CSR ("delta-8", include/sisua_b200.h: sisua_unpack_counts_csr8_u16) of a non-negative integer matrix [n, G] with
  counts below 65 536: (indptr int64 [n + 1] into ents, ents uint16, big_ptr int64 [n + 1] into big, big uint16)."""
  import numpy as np
  X = np.ascontiguousarray(X)
  n = X.shape[0]
  r, c = np.nonzero(X)
  v = X[r, c].astype(np.int64)
  first = np.ones(r.size, dtype=bool)
  first[1:] = r[1:] != r[:-1]
  prev = np.empty_like(c)
  prev[1:] = c[:-1]
  prev[first] = 0
  delta = (c - prev).astype(np.int64)                 # advance from the previous entry of the row (first: from column 0)
  skips = delta // 255                                # pure-skip words (advance 255, count 0) in front of the entry
  rem = delta - 255 * skips
  words_per = skips + 1
  tot = int(words_per.sum())
  ents = np.full(tot, 255, dtype=np.uint16)           # skip words by default: advance 255, count 0
  pos = np.cumsum(words_per) - 1                      # position of each real entry
  v8 = np.minimum(v, 255)
  ents[pos] = (rem | (v8 << 8)).astype(np.uint16)
  indptr = np.zeros(n + 1, dtype=np.int64)
  np.add.at(indptr, r + 1, words_per)
  indptr = np.cumsum(indptr)
  esc = v >= 255
  big = v[esc].astype(np.uint16)
  big_ptr = np.zeros(n + 1, dtype=np.int64)
  np.add.at(big_ptr, r[esc] + 1, 1)
  big_ptr = np.cumsum(big_ptr)
  return indptr, ents, big_ptr, big


def decode_csr8(indptr, ents, big_ptr, big, genes):
  """NumPy inverse of `encode_csr8` (tests, and the documentation of the format)."""
  import numpy as np
  n = len(indptr) - 1
  X = np.zeros((n, genes), dtype=np.float32)
  for i in range(n):
    col, k = 0, int(big_ptr[i])
    for w in ents[int(indptr[i]):int(indptr[i + 1])]:
      col += int(w) & 0xff
      val = int(w) >> 8
      if val == 255:
        val = int(big[k]); k += 1
      if val:
        X[i, col] = val
  return X


class Csr8Batch:
  """Pinned host packed-CSR form of one minibatch (see `encode_csr8`): 2 bytes per non-zero over PCIe."""

  def __init__(self, X):
    import numpy as np
    X = np.ascontiguousarray(X)
    if X.min() < 0 or X.max() >= 65536 or not np.array_equal(X, np.rint(X)):
      raise ValueError("Csr8Batch needs non-negative integer counts below 65536")
    ip, ents, bp, big = encode_csr8(X)
    self.rows, self.genes = X.shape
    self.indptr = _pin(torch.from_numpy(ip.astype(np.int32)))
    self.big_ptr = _pin(torch.from_numpy(bp.astype(np.int32)))
    self.ents = _pin(torch.from_numpy(ents.view(np.int16)))
    self.big = _pin(torch.from_numpy(np.concatenate([big, np.zeros(1, np.uint16)]).view(np.int16)))      # never empty

  @classmethod
  def view(cls, rows, genes, indptr, big_ptr, ents, big) -> "Csr8Batch":
    b = cls.__new__(cls)
    b.rows, b.genes, b.indptr, b.big_ptr, b.ents, b.big = rows, genes, indptr, big_ptr, ents, big
    return b

  @property
  def nbytes(self) -> int:
    return (self.indptr.numel() + self.big_ptr.numel()) * 4 + (self.ents.numel() + self.big.numel()) * 2


class HostDataset:
  """A training set kept in pinned HOST memory and served as minibatches for `HostTrainPipeline` (what `fit(data_on='host')`
  uses).  Mirrors the reference's input pipeline (sisua/data/_single_cell_base.py:593-601: map -> cache -> shuffle ->
  batch(drop_remainder)): the rows are permuted ONCE when the cache is built (the reference shuffles through a
  1000-element buffer) and every step takes the next B rows.  Integer count matrices are cached in CSR form (int64 row
  pointers, uint16 gene ids and counts; single-cell matrices are 70-96 % zeros), so a minibatch is three pinned slices
  plus B + 1 rebased row pointers; anything else stays dense float32."""

  def __init__(self, data, batch: int, shuffle: bool = True, seed: int = 0, with_y: bool = False, with_library: bool = False,
               packed: bool = False):
    """`packed`: cache integer counts in the 2-bytes-per-non-zero packed CSR form (`Csr8Batch`) instead of 4-byte CSR; only
    the CUDA-graph path of `HostTrainPipeline` consumes it."""
    import numpy as np
    X = np.asarray(data.X)
    N, G = X.shape
    self.B, self.genes, self.n = int(batch), G, N
    order = np.random.default_rng(seed).permutation(N) if shuffle else np.arange(N)
    self.order = order
    self.max_count = float(X.max()) if X.size else 0.0
    self.h2d_bytes, self.batches_served = 0, 0
    integer = G <= 65536 and X.min() >= 0 and self.max_count < 65536
    if integer:
      for s in range(0, N, 65536):           # integrality check in chunks (no full-size temporary)
        blk = X[s:s + 65536]
        if not np.array_equal(blk, np.rint(blk)):
          integer = False
          break
    self.csr = integer
    self.packed = bool(packed and integer)
    if self.packed:
      ips, ents_l, bps, big_l = [np.zeros(1, np.int64)], [], [np.zeros(1, np.int64)], []
      for s in range(0, N, 32768):
        ip, ents, bp, big = encode_csr8(X[order[s:s + 32768]])
        ips.append(ip[1:] + ips[-1][-1]); bps.append(bp[1:] + bps[-1][-1]); ents_l.append(ents); big_l.append(big)
      self.indptr, self.big_ptr = np.concatenate(ips), np.concatenate(bps)
      self.ents = _pin(torch.from_numpy(np.concatenate(ents_l).view(np.int16)))
      self.big = _pin(torch.from_numpy(np.concatenate(big_l + [np.zeros(1, np.uint16)]).view(np.int16)))
      self._ip = [_pin(torch.empty(self.B + 1, dtype=torch.int32)) for _ in range(kRing)]
      self._bp = [_pin(torch.empty(self.B + 1, dtype=torch.int32)) for _ in range(kRing)]
    elif integer:
      counts = np.zeros(N + 1, dtype=np.int64)
      cols_l, vals_l = [], []
      for s in range(0, N, 32768):
        blk = X[order[s:s + 32768]]
        r, c = np.nonzero(blk)
        counts[1 + s:1 + s + blk.shape[0]] = np.bincount(r, minlength=blk.shape[0])
        cols_l.append(c.astype(np.uint16)); vals_l.append(blk[r, c].astype(np.uint16))
      self.indptr = np.cumsum(counts)
      self.cols = _pin(torch.from_numpy(np.concatenate(cols_l).view(np.int16)))
      self.vals = _pin(torch.from_numpy(np.concatenate(vals_l).view(np.int16)))
      self._ip = [_pin(torch.empty(self.B + 1, dtype=torch.int32)) for _ in range(kRing)]
    else:
      self.dense = _pin(torch.from_numpy(np.ascontiguousarray(X[order], dtype=np.float32)))
    self.y = _pin(torch.from_numpy(np.ascontiguousarray(data.Y[order], dtype=np.float32))) if (with_y and data.Y is not None) else None
    self.mask = _pin(torch.from_numpy(np.ascontiguousarray(data.mask[order], dtype=np.uint8))) if with_y else None
    self.library = _pin(torch.from_numpy(np.ascontiguousarray(data.library[order], dtype=np.float32))) if with_library else None

  def __len__(self):
    return self.n // self.B

  def batch_at(self, s: int):
    """(x, extras): minibatch s of the cached order; x is a `CsrBatch` view or a pinned float32 [B, G] slice."""
    B = self.B
    lo, hi = s * B, (s + 1) * B
    if self.packed:
      a, b = int(self.indptr[lo]), int(self.indptr[hi])
      a2, b2 = int(self.big_ptr[lo]), int(self.big_ptr[hi])
      ip = self._ip[self.batches_served % len(self._ip)]; bp = self._bp[self.batches_served % len(self._bp)]
      ip.copy_(torch.from_numpy((self.indptr[lo:hi + 1] - a).astype("int32")))
      bp.copy_(torch.from_numpy((self.big_ptr[lo:hi + 1] - a2).astype("int32")))
      x = Csr8Batch.view(B, self.genes, ip, bp, self.ents[a:b], self.big[a2:max(b2, a2 + 1)])
      self.h2d_bytes += x.nbytes
    elif self.csr:
      a, b = int(self.indptr[lo]), int(self.indptr[hi])
      ip = self._ip[self.batches_served % len(self._ip)]
      ip.copy_(torch.from_numpy((self.indptr[lo:hi + 1] - a).astype("int32")))
      x = CsrBatch.view(B, self.genes, ip, self.cols[a:b], self.vals[a:b])
      self.h2d_bytes += x.nbytes
    else:
      x = self.dense[lo:hi]
      self.h2d_bytes += x.numel() * 4
    extras = {}
    if self.y is not None:
      extras["y"] = self.y[lo:hi]; extras["mask"] = self.mask[lo:hi]
      self.h2d_bytes += extras["y"].numel() * 4 + B
    if self.library is not None:
      extras["library"] = self.library[lo:hi]
      self.h2d_bytes += B * 8
    self.batches_served += 1
    return x, extras

  def batch(self, epoch: int, s: int):
    return self.batch_at(s % max(1, len(self)))


def _capture_step(eng: Engine, run: Callable[[], None]) -> "torch.cuda.CUDAGraph":
  """Warm `run` up once outside capture (restoring the optimiser state it touches), then capture it."""
  dev = eng.device
  side = torch.cuda.Stream(device=dev)
  side.wait_stream(torch.cuda.current_stream(dev))
  with torch.cuda.stream(side):
    snap = [t.clone() for t in (eng.params, eng.adam_m, eng.adam_v, eng.bn_moving)]
    step0 = eng.step_count
    run()
    side.synchronize()
    for t, s_ in zip((eng.params, eng.adam_m, eng.adam_v, eng.bn_moving), snap):
      t.copy_(s_)
    eng.reset_step_counter(step0)
  torch.cuda.current_stream(dev).wait_stream(side)
  graph = torch.cuda.CUDAGraph()
  with torch.cuda.graph(graph):
    run()
  eng.reset_step_counter(step0)            # capture does not execute, but keep host / device counters aligned
  return graph


class HostTrainPipeline:
  """Train from pinned host minibatches (float32 / 16-bit dense or `CsrBatch`).

  Single GPU: each of `depth` slots owns static device staging buffers and one CUDA graph
  (widen counts -> train step -> Adam -> loss D2H); `step` enqueues the H2D copies on a side stream and replays the
  slot's graph, so the transfer of step i+1 runs under the kernels of step i.  The graph matters here: with eager
  launches the per-kernel command fetches queue behind the bulk H2D traffic on PCIe and the ~20 short kernels of a
  step each start late (measured: 417 -> 700 us per step under a concurrent 11 MB copy; graph replay: 411 -> 423 us).
  Data parallel: the slot owns two graphs (… -> train step | Adam -> loss D2H) and the `allreduce(grads)` callback is
  launched eagerly between them, so only the collective's own launches see the busy PCIe link.  Ragged batches and
  model families with extra host inputs go through the library's host-buffer entry point `sisua_train_step_host`,
  which stages on its own copy stream."""

  def __init__(self, eng: Engine, batch: int, depth: int = 2, use_graph: bool = True):
    assert 1 <= depth < kRing
    self.eng = eng
    self.batch = batch
    self.depth = depth
    self.use_graph = use_graph
    dev, cfg = eng.device, eng.cfg
    self.copy_stream = torch.cuda.Stream(device=dev)
    self.host_loss = [torch.empty((1,), dtype=torch.float32).pin_memory() for _ in range(max(depth, 4))]
    self.slots = [None] * depth
    self.graphs = {}
    self.i = 0
    self.last_loss_dev = None      # device scalar of the most recent graph-replayed step (None after an eager step)

  def _slot(self, s: int):
    if self.slots[s] is None:
      eng, B = self.eng, self.batch
      dev, G = eng.device, eng.cfg.n_genes
      self.slots[s] = dict(
          x=None, x16=None, csr=None, csr8=None,
          eps=torch.zeros((B, eng.cfg.n_latent), device=dev),
          terms=torch.empty((5, B), device=dev), loss=torch.empty((1,), device=dev),
          filled=torch.cuda.Event(), consumed=torch.cuda.Event())
      self.slots[s]["consumed"].record()
    return self.slots[s]

  def _graph(self, s: int, fmt: str, lr: float, clipnorm: float, world: int = 1, split: bool = False, seed: int = 0,
             with_eps: bool = True, peer: bool = False):
    """(graph, None), or (backward graph, optimiser graph) when an all-reduce sits between them."""
    key = (s, fmt, float(lr), float(clipnorm), int(world), bool(split), int(seed), bool(with_eps), bool(peer))
    if key not in self.graphs:
      eng, sl = self.eng, self._slot(s)
      def fwd_bwd():
        # integer formats stay 16-bit on the device: CSR was decoded into the slot's uint16 matrix on the COPY stream right
        # behind its H2D transfer (`step`), i.e. under the kernels of the previous step and off this graph's critical path;
        # a dense uint16 batch is used as it arrived; the step's streaming kernels widen the counts themselves
        # (sisua_train_step_gather_u16)
        x_dev = sl["x"] if fmt == "f32" else sl["x16"]
        # eps from the host when one is shipped (tests), otherwise Philox noise drawn in-kernel
        eng.train_step(x_dev, eps_z=sl["eps"] if (with_eps and eng.cfg.model_kind != 2) else None, terms=sl["terms"],
                       loss=sl["loss"], seed=seed, step=-1)
      def optimise():
        if peer:       # data parallel over peer memory: the exchange is part of the optimiser kernel (and of the graph)
          eng.adam_step_dp(lr=lr, clipnorm=clipnorm, t=0)
        else:
          eng.adam_step(lr=lr, clipnorm=clipnorm, grad_scale=1.0 / world, t=0)
        self.host_loss[s].copy_(sl["loss"], non_blocking=True)
      if not split:
        self.graphs[key] = (_capture_step(eng, lambda: (fwd_bwd(), optimise())), None)
      else:
        # captured separately; the warm-up of the first one leaves gradients behind that the second one's warm-up
        # consumes, and _capture_step restores parameters and optimiser state after each
        self.graphs[key] = (_capture_step(eng, fwd_bwd), _capture_step(eng, optimise))
    return self.graphs[key]

  def step(self, x_host, eps_host: Optional[torch.Tensor], step: int, lr: float = 1e-3, clipnorm: float = 100.0,
           world: int = 1, allreduce: Optional[Callable] = None, seed: int = 0, peer=None, **host_extras):
    """Enqueue one train step; returns the pinned host tensor that will hold the loss once the stream has drained
    (read it after `flush`; it is reused `depth` steps later)."""
    eng = self.eng
    is_csr8 = isinstance(x_host, Csr8Batch)
    is_csr = isinstance(x_host, CsrBatch)
    rows = x_host.rows if (is_csr or is_csr8) else x_host.shape[0]
    graphable = (self.use_graph and not host_extras and rows == self.batch and
                 eng.cfg.model_kind in (0, 2) and eng.cfg.n_proteins == 0)
    if not graphable:
      if is_csr8:
        raise ValueError("packed CSR minibatches are only consumed by the CUDA-graph path (plain VAE / DCA, full batches)")
      out = self.host_loss[self.i % len(self.host_loss)]
      self.i += 1
      self.last_loss_dev = None
      eng.train_step_host(x_host, eps_z=eps_host, host_loss=out, seed=seed, step=step, **host_extras)
      if peer is not None:
        eng.adam_step_dp(lr=lr, clipnorm=clipnorm, t=step)
        return out
      if allreduce is not None:
        allreduce(eng.grads)
      eng.adam_step(lr=lr, clipnorm=clipnorm, grad_scale=1.0 / world, t=step)
      return out
    s = self.i % self.depth
    self.i += 1
    sl = self._slot(s)
    fmt = "csr8" if is_csr8 else ("csr" if is_csr else ("u16" if x_host.dtype in (torch.int16, torch.uint16) else "f32"))
    if fmt == "csr8" and sl.get("csr8") is None:      # worst-case capacities: the graph bakes the addresses in
      cap = self.batch * eng.cfg.n_genes
      i32 = lambda n: torch.zeros(n, device=eng.device, dtype=torch.int32)
      i16 = lambda n: torch.zeros(n, device=eng.device, dtype=torch.int16)
      sl["csr8"] = (i32(self.batch + 1), i32(self.batch + 1), i16(cap + cap // 255 + 8), i16(cap + 8))
    if fmt == "csr" and sl["csr"] is None:      # worst-case capacity: the graph bakes the addresses in
      cap = self.batch * eng.cfg.n_genes
      sl["csr"] = (torch.zeros(self.batch + 1, device=eng.device, dtype=torch.int32),
                   torch.zeros(cap, device=eng.device, dtype=torch.int16), torch.zeros(cap, device=eng.device, dtype=torch.int16))
    if fmt == "f32" and sl["x"] is None:
      sl["x"] = torch.empty((self.batch, eng.cfg.n_genes), device=eng.device)
    if fmt in ("u16", "csr", "csr8") and sl["x16"] is None:
      sl["x16"] = torch.zeros((self.batch, eng.cfg.n_genes), device=eng.device, dtype=torch.int16)
    graph, graph_opt = self._graph(s, fmt, lr, clipnorm, world, allreduce is not None, seed, eps_host is not None, peer is not None)
    if eng.step_count != step - 1:             # the graph follows the device-side step counter (dropout masks, Adam t)
      eng.reset_step_counter(step - 1)
    main = torch.cuda.current_stream(eng.device)
    # Host-side throttle: the H2D copies of this slot's previous use must have been DONE before more batches are prepared.
    # The dataset hands out its rebased row-pointer arrays from a small pinned ring; without the wait a host running many
    # steps ahead of the copy engine would overwrite a ring entry whose DMA has not happened yet.  (`depth` steps of
    # run-ahead remain, which is what hides the host's per-step cost.)
    if sl.get("used"):
      sl["filled"].synchronize()
    sl["used"] = True
    with torch.cuda.stream(self.copy_stream):
      self.copy_stream.wait_event(sl["consumed"])
      if fmt == "csr8":
        ip, bp, ee, bb = sl["csr8"]
        ip.copy_(x_host.indptr, non_blocking=True); bp.copy_(x_host.big_ptr, non_blocking=True)
        ee[:x_host.ents.numel()].copy_(x_host.ents, non_blocking=True)
        bb[:x_host.big.numel()].copy_(x_host.big, non_blocking=True)
        eng.unpack_counts_csr8(ip, bp, ee, bb, sl["x16"])        # decode on the copy stream, behind the transfer
      elif fmt == "csr":
        nnz = x_host.cols.numel()
        ip, cc, vv = sl["csr"]
        ip.copy_(x_host.indptr, non_blocking=True)
        cc[:nnz].copy_(x_host.cols, non_blocking=True)
        vv[:nnz].copy_(x_host.vals, non_blocking=True)
        eng.unpack_counts_csr(ip, cc, vv, sl["x16"])
      elif fmt == "u16":
        sl["x16"].copy_(x_host.view(torch.int16), non_blocking=True)
      else:
        sl["x"].copy_(x_host, non_blocking=True)
      if eps_host is not None:
        sl["eps"].copy_(eps_host, non_blocking=True)
      sl["filled"].record(self.copy_stream)
    main.wait_event(sl["filled"])
    graph.replay()
    sl["consumed"].record(main)
    if graph_opt is not None:
      allreduce(eng.grads)
      graph_opt.replay()
    eng.step_count += 1
    self.last_loss_dev = sl["loss"]
    return self.host_loss[s]

  def flush(self, losses: List[torch.Tensor]):
    torch.cuda.current_stream(self.eng.device).synchronize()
    return [float(l.item()) for l in losses[-self.depth:]]

  def flush_all(self, losses: List[torch.Tensor]):
    """Synchronise and return the losses that are still readable: the pinned result slots are reused every `depth`
    steps, so only the last `depth` entries of `losses` are distinct values."""
    return self.flush(losses)


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
    self.graph = _capture_step(eng, lambda: self._run(lr, clipnorm, seed))

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


class GraphedGatherStep:
  """`sisua_train_step_gather` + `sisua_adam_step` captured once and replayed per minibatch: the step reads its rows out
  of the HBM-resident matrices through a static index buffer, draws dropout masks and reparameterisation noise from the
  device-side step counter, so a replay only needs the B row indices copied in (what `fit()` does for minibatches
  <= 2048, where the ~20 kernels of a step are launch-bound)."""

  def __init__(self, eng: Engine, batch: int, x_all: torch.Tensor, y_all=None, library_all=None, mask_all=None,
               lr: float = 1e-3, clipnorm: float = 100.0, seed: int = 0):
    dev = eng.device
    self.eng, self.batch = eng, batch
    self.rows = torch.arange(batch, device=dev, dtype=torch.int32)
    self.terms = torch.empty((5, batch), device=dev)
    self.loss = torch.empty((1,), device=dev)
    self._keep = (x_all, y_all, library_all, mask_all)

    def run():
      eng.train_step_gather(x_all, self.rows, y_all=y_all, library_all=library_all, mask_all=mask_all, terms=self.terms,
                            loss=self.loss, seed=seed, step=-1)
      eng.adam_step(lr=lr, clipnorm=clipnorm, grad_scale=1.0, t=0)
    self.graph = _capture_step(eng, run)

  def step(self, rows: torch.Tensor):
    self.rows.copy_(rows, non_blocking=True)
    self.graph.replay()
    self.eng.step_count += 1
    return self.terms, self.loss
