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
