"""Data-parallel plumbing: cells are sharded across the GPUs of one box, one process per GPU; the only
exchange per step is a sum all-reduce of the flat fp32 gradient buffer (NCCL over NVLink), scaled by
1/world inside the fused Adam kernel.  BatchNorm statistics stay local to a rank
during training (what tf.distribute would do); the moving statistics are averaged when training ends so
every replica serves identical inference results."""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(n_cells: int, rank: int, world: int) -> Tuple[int, int]:
  """Contiguous, balanced [begin, end) slice of the cells owned by `rank` (sizes differ by at most one)."""
  if not (0 <= rank < world):
    raise ValueError(f"rank {rank} outside world {world}")
  base, rem = divmod(n_cells, world)
  begin = rank * base + min(rank, rem)
  return begin, begin + base + (1 if rank < rem else 0)


def steps_per_epoch(n_cells: int, world: int, batch_per_rank: int) -> int:
  """All ranks must take the same number of steps: limited by the smallest shard (drop_remainder=True)."""
  smallest = n_cells // world
  return smallest // batch_per_rank


def allreduce_gradients(flat_grads: torch.Tensor) -> float:
  """Sum-all-reduce the flat gradient buffer in place; returns the scale (1/world) the optimiser must apply."""
  if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
    dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM)
    return 1.0 / dist.get_world_size()
  return 1.0


def average_moving_statistics(bn_moving: torch.Tensor) -> None:
  if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
    dist.all_reduce(bn_moving, op=dist.ReduceOp.SUM)
    bn_moving.div_(dist.get_world_size())


def broadcast_parameters(*tensors: torch.Tensor, src: int = 0) -> None:
  if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
    for t in tensors:
      dist.broadcast(t, src=src)


class OverlappedAllReduce:
  """Two-bucket gradient all-reduce: bucket 0 (output heads, ~3/4 of the bytes) starts on a side stream as soon as
  the step signals it is final and runs under the rest of the backward pass; bucket 1 (everything else) follows
  the step.  Same collective order on every rank (bucket 0 then bucket 1)."""

  def __init__(self, engine):
    self.eng = engine
    self.enabled = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    if self.enabled:
      self.event, (self.b0, self.e0) = engine.enable_grad_overlap()
      self.side = torch.cuda.Stream(device=engine.device)

  def __call__(self) -> float:
    """Call right after train_step has been enqueued. Returns the gradient scale for adam_step."""
    if not self.enabled:
      return 1.0
    g = self.eng.grads
    main = torch.cuda.current_stream(self.eng.device)
    self.side.wait_event(self.event)
    with torch.cuda.stream(self.side):
      w0 = dist.all_reduce(g[self.b0:self.e0], async_op=True)
    w1 = dist.all_reduce(g[:self.b0], async_op=True)
    w2 = dist.all_reduce(g[self.e0:], async_op=True) if self.e0 < g.numel() else None   # empty: the heads are laid out last
    w0.wait(); w1.wait()
    if w2 is not None:
      w2.wait()
    main.wait_stream(self.side)
    return 1.0 / dist.get_world_size()


class PeerExchange:
  """Gradient exchange + optimiser as ONE kernel over NVLink peer memory (csrc/dp_adam.cuh, `sisua_adam_step_dp`):
  reduce-scatter by peer loads -> per-variable clipnorm -> Adam on this rank's shard -> all-gather by peer stores.

  torch is only the plumbing: `torch.distributed._symmetric_memory` allocates the four symmetric buffers (flat gradients,
  flat parameters, the partial-norm table, the flag words) and exchanges the peer mappings; the engine's parameter and
  gradient buffers are moved into them.  After this, a data-parallel step is `train_step` + `adam_step_dp` -- no NCCL call
  on the data path, and the whole step can be captured in one CUDA graph."""

  def __init__(self, engine, group=None, grid: int = 0):
    import torch.distributed._symmetric_memory as symm
    if not (dist.is_available() and dist.is_initialized()):
      raise RuntimeError("PeerExchange needs an initialised torch.distributed process group")
    group = group or dist.group.WORLD
    self.eng, self.world, self.rank = engine, dist.get_world_size(group), dist.get_rank(group)
    if self.world > 8:
      raise ValueError("the peer-memory optimiser step supports up to 8 ranks (one NVSwitch box)")
    dev, total = engine.device, engine.total
    with torch.cuda.device(dev):
      self.params = symm.empty(total, dtype=torch.float32, device=dev)
      self.grads = symm.empty(total, dtype=torch.float32, device=dev)
      self.sq = symm.empty(8 * 48, dtype=torch.float64, device=dev)
      self.flags = symm.empty(64, dtype=torch.int32, device=dev)
      self.sq.zero_(); self.flags.zero_()
      handles = [symm.rendezvous(t, group.group_name) for t in (self.grads, self.params, self.sq, self.flags)]
    ptrs = []
    for t, hd in zip((self.grads, self.params, self.sq, self.flags), handles):
      off = int(getattr(hd, "offset", 0) or 0)
      p = [int(b) + off for b in hd.buffer_ptrs]
      if p[self.rank] != t.data_ptr():
        raise RuntimeError("symmetric-memory handle does not describe the tensor it was made for")
      ptrs.append(p)
    self._handles = handles
    engine.rebind(self.params, self.grads)
    torch.cuda.synchronize(dev)
    dist.barrier(group)                      # every rank's buffers are initialised before anybody's kernel touches them
    engine.dp_bind(self.rank, self.world, *ptrs, grid=grid)
    self.shard = engine.dp_shard()

  def step(self, lr=1e-3, clipnorm=100.0, t=0):
    self.eng.adam_step_dp(lr=lr, clipnorm=clipnorm, t=t)

  def gather_optimizer_state(self):
    """Adam moments are maintained per shard; assemble the full vectors on every rank (checkpointing)."""
    b, e = self.shard
    for t in (self.eng.adam_m, self.eng.adam_v):
      t[:b].zero_(); t[e:].zero_()
      dist.all_reduce(t, op=dist.ReduceOp.SUM)
