"""Parity at the launch geometry bench.py actually measures.

At 18 944 cells x 2 000 genes every fused output-head CTA walks all 63 gene tiles in ONE chunk (each mbarrier stage
is recycled 31 times) and the first-layer kernel runs one 32-k-block chunk without atomics; none of the small-batch
parity tests reaches that depth.  These tests compare the per-cell ELBO, the loss and ALL gradients with the fp64
oracle at those shapes, and assert the walk depth through sisua_debug_geometry so a heuristic change cannot silently
shrink what they cover.  Tolerances: BASELINE.json north_star (1e-4 relative on per-cell ELBO); gradients: the fused
path's gradient GEMMs run on single fp16 operands (2^-11 relative), bound 6e-3 of each tensor's largest entry."""
import numpy as np
import pytest
import torch

from oracle import step_oracle as O
from sisua_b200 import config as C
from sisua_b200 import params as PR
from tests import helpers as Hh

pytestmark = pytest.mark.gpu


def _close(a, b, rtol, atol=0.0, what=""):
  a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
  err = np.abs(a - b)
  bad = err > rtol * np.abs(b) + atol
  assert not bad.any(), f"{what}: {bad.sum()} / {bad.size} off; worst rel {np.max(err / (np.abs(b) + 1e-30)):.3e} abs {err.max():.3e}"


def _full_check(cfg, B, seed=11, step=3, gtol=6e-3, force=None, expect=None):
  from sisua_b200.engine import Engine
  flat = Hh.randomize_norm_params(cfg, PR.init_flat_params(cfg))
  mov = PR.init_bn_moving(cfg)
  batch = Hh.make_batch(cfg, B, seed=1)
  drop = Hh.oracle_dropout_masks(cfg, B, seed=seed, step=step)
  # no ReLU input of this batch within float32 rounding of zero (see helpers.separate_relu_ties for why)
  margin = Hh.separate_relu_ties(cfg, flat, mov, batch, drop)
  assert margin > 2e-6, margin
  eng = Engine(cfg, 0, flat_params=flat, bn_moving=mov)
  if force:
    eng.force_chunks(**force)
  geo = eng.geometry(B)
  for k, v in (expect or {}).items():
    assert geo[k] == v, f"geometry {k} = {geo[k]}, the test is written for {v}: {geo}"
  terms, loss = eng.train_step(seed=seed, step=step, **batch)
  torch.cuda.synchronize()
  P = Hh.oracle_params(cfg, flat)
  for p in P.values():
    p.requires_grad_(True)
  ref = O.forward(cfg, P, Hh.oracle_moving(cfg, mov), training=True, drop=drop, **batch)
  ref["loss"].backward()
  _close(terms[0].cpu().numpy(), ref["elbo"].detach().numpy(), 1e-4, what="per-cell ELBO")
  _close(terms[1].cpu().numpy(), ref["llk_x"].detach().numpy(), 1e-4, what="per-cell llk_x")
  _close(loss.cpu().numpy()[0], float(ref["loss"]), 1e-5, what="loss")
  got = eng.grads_dict()
  worst, cosines = {}, {}
  for name, p in P.items():
    g_ref = p.grad.numpy() if p.grad is not None else np.zeros(p.shape)
    scale = np.abs(g_ref).max() + 1e-12
    worst[name] = float(np.abs(got[name] - g_ref).max() / scale)
    # direction of the whole tensor: fp16-grade element noise must not tilt it
    cosines[name] = float((got[name].ravel().astype(np.float64) @ g_ref.ravel()) /
                          (np.linalg.norm(got[name].ravel().astype(np.float64)) * np.linalg.norm(g_ref.ravel()) + 1e-300))
  report = "; ".join(f"{k}: err/max {worst[k]:.2e} cos {cosines[k]:.6f}" for k in worst)
  print("gradient parity:", report)
  assert max(worst.values()) <= gtol, report
  assert min(cosines.values()) >= 0.9999, report
  eng.close()
  return geo, worst


def test_bench_geometry_vae_18944x2000_all_gradients():
  """bench.py's default shape: vae / zinbd, 18 944 cells x 2 000 genes, input dropout 0.3."""
  B, G = 18944, 2000
  cfg = C.make_step_config("vae", n_genes=G, n_latent=10, max_batch=B, input_dropout=0.3)
  geo, _ = _full_check(cfg, B)
  if torch.cuda.get_device_properties(0).multi_processor_count == 148:
    assert geo["out_chunks"] == 1 and geo["out_tiles_per_chunk"] == 63, geo       # the 63-tile single-chunk walk
    assert geo["enc_chunks"] == 1 and geo["enc_kblocks_per_chunk"] == 32, geo     # no split-K atomics
    assert geo["out_cell_tiles"] == 148


def test_bench_geometry_scvi_4736x2000_all_gradients():
  """scVI (three passes over the gene tiles, gene softmax across 16 chunks) at a multi-tile-per-chunk geometry."""
  B, G = 4736, 2000
  cfg = C.make_step_config("scvi", n_genes=G, n_latent=10, max_batch=B, enc_dropout=0.1, encl_dropout=0.1)
  geo, _ = _full_check(cfg, B)
  assert geo["out_tiles_per_chunk"] >= 12, geo


def test_config4_dca_5000_genes_all_gradients():
  """BASELINE.json configs[3] (dca, 5 000 genes): 157 gene tiles, 79 first-layer k-blocks."""
  B, G = 2048, 5000
  cfg = C.make_step_config("dca", n_genes=G, n_latent=10, max_batch=B, input_dropout=0.3)
  geo, _ = _full_check(cfg, B)
  assert geo["out_tiles_per_chunk"] >= 16, geo


@pytest.mark.parametrize("model,kw", [("vae", {}), ("scvi", {}), ("sisua", dict(n_proteins=10))])
def test_single_chunk_deep_walk_small_batch(model, kw):
  """The deep single-chunk walk (every stage recycled ~31 times, accumulator / gradient-tile / weight-gradient stages
  included) forced at a small batch for every model family, gradients included."""
  B, G = 300, 2000
  cfg = C.make_step_config(model, n_genes=G, n_latent=10, max_batch=B, input_dropout=0.3, **kw)
  _full_check(cfg, B, force=dict(out_chunks=1, enc_chunks=1, bwd_chunks=1),
              expect=dict(out_chunks=1, out_tiles_per_chunk=63, enc_chunks=1, enc_kblocks_per_chunk=32, bwd_chunks=1))
