"""world_size-2 gloo tests (CPU) of the data-parallel host logic: shard arithmetic, gradient averaging and
its equivalence with the single-process global batch (oracle gradients; no BatchNorm so shards do not couple)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sisua_b200 import distributed as DP


def test_shard_range_partitions_everything():
  for n, w in [(10, 2), (1_000_000, 8), (7, 3), (8381, 4)]:
    spans = [DP.shard_range(n, r, w) for r in range(w)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))
    sizes = [e - b for b, e in spans]
    assert max(sizes) - min(sizes) <= 1
  assert DP.steps_per_epoch(1_000_000, 8, 8192) == 15
  with pytest.raises(ValueError):
    DP.shard_range(10, 2, 2)


def _free_port():
  s = socket.socket()
  s.bind(("127.0.0.1", 0))
  p = s.getsockname()[1]
  s.close()
  return p


def _worker(rank, world, port, out):
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  dist.init_process_group("gloo", rank=rank, world_size=world)
  try:
    from oracle import step_oracle as O
    from sisua_b200 import config as C
    from sisua_b200 import params as PR
    from tests import helpers as Hh
    torch.set_num_threads(1)
    cfg = C.make_step_config("sisua", n_genes=40, n_proteins=4, n_latent=5, batchnorm=False)
    flat = Hh.randomize_norm_params(cfg, PR.init_flat_params(cfg))
    full = Hh.make_batch(cfg, 32, seed=3)
    b, e = DP.shard_range(32, rank, world)
    mine = {k: v[b:e] for k, v in full.items()}

    def grads_of(batch):
      P = Hh.oracle_params(cfg, flat)
      for p in P.values():
        p.requires_grad_(True)
      O.forward(cfg, P, None, training=True, **batch)["loss"].backward()
      return torch.cat([p.grad.reshape(-1) for p in P.values()])

    g = grads_of(mine).float()
    scale = DP.allreduce_gradients(g)
    g = g * scale
    moving = torch.full((2, 2, 4), float(rank))
    DP.average_moving_statistics(moving)
    p = torch.full((3,), float(rank + 5))
    DP.broadcast_parameters(p, src=0)
    if rank == 0:
      ref = grads_of(full).float()
      out.put((float((g - ref).abs().max()), float(ref.abs().max()), scale, float(moving.mean()), float(p[0])))
    else:
      out.put(("p", float(p[0])))
  finally:
    dist.destroy_process_group()


def test_two_rank_gradient_average_equals_global_batch():
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  port = _free_port()
  procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
  for p in procs:
    p.start()
  results = [q.get(timeout=180) for _ in range(2)]
  for p in procs:
    p.join(timeout=60)
    assert p.exitcode == 0
  main = [r for r in results if r[0] != "p"][0]
  other = [r for r in results if r[0] == "p"][0]
  err, scale_ref, scale, moving_mean, p0 = main
  assert scale == 0.5
  assert err <= 1e-5 * max(1.0, scale_ref), (err, scale_ref)
  assert moving_mean == 0.5          # (0 + 1) / 2
  assert p0 == 5.0 and other[1] == 5.0   # broadcast from rank 0
