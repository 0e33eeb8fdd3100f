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
