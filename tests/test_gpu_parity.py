"""CUDA path vs the CPU oracle on identical synthetic counts, weights and injected eps.
Tolerance (BASELINE.json north_star): per-cell ELBO, latent means and imputed means within 1e-4
relative in fp32 (gemm_mode 0 = exact fp32 FFMA, gemm_mode 1 = error-compensated 3xTF32)."""
import numpy as np
import pytest
import torch

from oracle import step_oracle as O
from sisua_b200 import config as C
from sisua_b200 import params as PR
from tests import helpers as Hh

pytestmark = pytest.mark.gpu

RTOL = 1e-4


def _engine(cfg, flat, mov):
  from sisua_b200.engine import Engine
  return Engine(cfg, 0, flat_params=flat, bn_moving=mov)


def _close(a, b, rtol=RTOL, atol=0.0, what=""):
  a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
  err = np.abs(a - b)
  tol = rtol * np.abs(b) + atol
  bad = err > tol
  assert not bad.any(), f"{what}: {bad.sum()} / {bad.size} off; worst rel {np.max(err / (np.abs(b) + 1e-30)):.3e} abs {err.max():.3e}"


MODELS = [("vae", {}), ("scvi", {}), ("dca", {}), ("sisua", dict(n_proteins=10))]
MODES = [C.GEMM_FP32_UNFUSED, C.GEMM_TC_3XFP16]


def _setup(model, kw, G, B, mode, seed=0, trained_moving=True, **cfgkw):
  cfg = C.make_step_config(model, n_genes=G, gemm_mode=mode, max_batch=max(B * 4, 256), **kw, **cfgkw)
  flat = Hh.randomize_norm_params(cfg, PR.init_flat_params(cfg))
  mov = PR.init_bn_moving(cfg)
  if trained_moving:
    rng = np.random.default_rng(5)
    mov[:, 0, :] = rng.normal(0, 0.3, mov[:, 0, :].shape)
    mov[:, 1, :] = rng.uniform(0.5, 2.0, mov[:, 1, :].shape)
  batch = Hh.make_batch(cfg, B, seed=seed)
  return cfg, flat, mov, batch


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("model,kw", MODELS)
@pytest.mark.parametrize("G,B", [(558, 96), (50, 33)])
def test_inference_parity(model, kw, G, B, mode):
  cfg, flat, mov, batch = _setup(model, kw, G, B, mode)
  eng = _engine(cfg, flat, mov)
  out = eng.infer(want_mean=True, want_disp=True, want_pi=True, **batch)
  torch.cuda.synchronize()
  ref = O.forward(cfg, Hh.oracle_params(cfg, flat), Hh.oracle_moving(cfg, mov), training=False, **batch)
  terms = out["terms"].cpu().numpy()
  _close(terms[0], ref["elbo"].numpy(), what="elbo")
  _close(terms[1], ref["llk_x"].numpy(), what="llk_x")
  _close(terms[3], ref["kl_z"].numpy(), atol=1e-5, what="kl_z")
  _close(out["z_loc"].cpu().numpy(), ref["z_loc"].numpy(), atol=1e-5, what="z_loc")
  _close(out["z_scale"].cpu().numpy(), ref["z_scale"].numpy(), atol=1e-6, what="z_scale")
  _close(out["mean"].cpu().numpy(), ref["mu"].numpy(), atol=1e-7, what="imputed mean")
  _close(out["disp"].cpu().numpy(), ref["theta"].numpy(), atol=1e-7, what="dispersion")
  if model == "sisua":
    _close(terms[2], ref["llk_y"].numpy(), what="llk_y")
    _close(out["y_mean"].cpu().numpy(), ref["y_mean"].numpy(), what="y_mean")
  if model == "scvi":
    _close(terms[4], ref["kl_l"].numpy(), atol=1e-5, what="kl_l")
  eng.close()


@pytest.mark.parametrize("mode", MODES)
def test_inference_mc_samples(mode):
  cfg, flat, mov, _ = _setup("vae", {}, 200, 40, mode)
  batch = Hh.make_batch(cfg, 40, S=3)
  eng = _engine(cfg, flat, mov)
  out = eng.infer(S=3, **batch)
  ref = O.forward(cfg, Hh.oracle_params(cfg, flat), Hh.oracle_moving(cfg, mov), training=False, **batch)
  _close(out["terms"][0].cpu().numpy().reshape(3, 40), ref["elbo"].numpy(), what="elbo[S,B]")
  _close(out["mean"].cpu().numpy().reshape(3, 40, 200), ref["mu"].numpy(), atol=1e-7, what="mean[S,B,G]")
  eng.close()


def _grad_check(cfg, flat, mov, batch, eng, gtol=None, drop=None, seed=0, step=-1):
  # fused path: the two gradient GEMMs of the output heads run on single fp16 operands (2^-11 relative)
  gtol = gtol or (2e-3 if cfg.gemm_mode == C.GEMM_FP32_UNFUSED else 6e-3)
  terms, loss = eng.train_step(seed=seed, step=step, **batch)
  torch.cuda.synchronize()
  P = Hh.oracle_params(cfg, flat)
  for p in P.values():
    p.requires_grad_(True)
  om = Hh.oracle_moving(cfg, mov)
  ref = O.forward(cfg, P, om, training=True, drop=drop, **batch)
  ref["loss"].backward()
  _close(terms[0].cpu().numpy(), ref["elbo"].detach().numpy(), what="train elbo")
  _close(loss.cpu().numpy()[0], float(ref["loss"]), what="loss")
  got = eng.grads_dict()
  for name, p in P.items():
    g_ref = p.grad.numpy() if p.grad is not None else np.zeros(p.shape)
    scale = np.abs(g_ref).max() + 1e-12
    err = np.abs(got[name] - g_ref).max()
    assert err <= gtol * scale + 1e-9, f"grad {name}: max err {err:.3e} vs scale {scale:.3e}"
  # BN moving statistics
  new_mov = PR.moving_to_dict(cfg, eng.bn_moving.cpu().numpy())
  for k, v in ref["new_moving"].items():
    _close(new_mov[k], v.numpy(), rtol=1e-4, atol=1e-6, what=k)


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("model,kw", MODELS)
@pytest.mark.parametrize("G,B", [(558, 128), (50, 33)])
def test_train_step_gradients(model, kw, G, B, mode):
  cfg, flat, mov, batch = _setup(model, kw, G, B, mode, trained_moving=False)
  eng = _engine(cfg, flat, mov)
  _grad_check(cfg, flat, mov, batch, eng)
  eng.close()


@pytest.mark.parametrize("mode", MODES)
def test_variants_nbd_nobn_stress(mode):
  # 'nbd' likelihood, no batch norm (bias path), dense stress counts (tests/test_scalability.py recipe)
  cfg = C.make_step_config("sisua", n_genes=120, n_proteins=10, x_dist="nbd", y_dist="nbd", batchnorm=False,
                           gemm_mode=mode, max_batch=256)
  flat = Hh.randomize_norm_params(cfg, PR.init_flat_params(cfg))
  mov = PR.init_bn_moving(cfg)
  batch = Hh.make_batch(cfg, 64, stress=True)
  eng = _engine(cfg, flat, mov)
  _grad_check(cfg, flat, mov, batch, eng)
  eng.close()


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("model,kw", [("vae", {}), ("scvi", {}), ("sisua", dict(n_proteins=10))])
def test_multi_step_training_matches_oracle(model, kw, mode):
  """T Adam steps on the GPU vs T oracle steps.  Adam's first updates are sign-like (|update| ~ lr whatever |g| is), so a
  bound on the parameter DIFFERENCE cannot fail below T * lr; what carries signal is the loss trajectory and the
  DIRECTION of the total update: cosine similarity per tensor and overall, plus a T = 1 check scaled to the update."""
  G, B, T = 300, 64, 5
  cfg, flat, mov, _ = _setup(model, kw, G, B, mode, trained_moving=False)
  eng = _engine(cfg, flat, mov)
  P = Hh.oracle_params(cfg, flat)
  P0 = {k: v.clone() for k, v in P.items()}
  om = Hh.oracle_moving(cfg, mov)
  m = {k: torch.zeros_like(v) for k, v in P.items()}
  v = {k: torch.zeros_like(p) for k, p in P.items()}

  def cosines():
    got = eng.params_dict()
    dg = {k: got[k].astype(np.float64) - P0[k].numpy() for k in P}
    dr = {k: P[k].numpy() - P0[k].numpy() for k in P}
    per = {k: float((dg[k].ravel() @ dr[k].ravel()) / (np.linalg.norm(dg[k]) * np.linalg.norm(dr[k]) + 1e-300)) for k in P}
    a = np.concatenate([dg[k].ravel() for k in P]); b = np.concatenate([dr[k].ravel() for k in P])
    return per, float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b)))

  for t in range(1, T + 1):
    batch = Hh.make_batch(cfg, B, seed=t)
    terms, loss = eng.train_step(**batch)
    eng.adam_step(lr=1e-3, clipnorm=100.0, t=t)
    out, _ = O.train_step(cfg, P, om, m, v, t, batch, lr=1e-3, clipnorm=100.0)
    _close(loss.cpu().numpy()[0], float(out["loss"]), rtol=2e-4, what=f"loss step {t}")
    if t == 1:
      # one step: update = lr * sign-like(g); entries whose gradient is below the arithmetic's noise may flip, the rest
      # must agree to a small fraction of the step
      got = eng.params_dict()
      flips, total = 0, 0
      for name, p in P.items():
        d = np.abs(got[name] - p.numpy())
        flips += int((d > 2e-4).sum()); total += d.size
      frac = flips / total
      assert frac <= (1e-4 if mode == C.GEMM_FP32_UNFUSED else 2e-2), f"{flips}/{total} entries moved differently in step 1"
  per, overall = cosines()
  lim = 0.9999 if mode == C.GEMM_FP32_UNFUSED else 0.999
  assert overall >= lim, f"total update direction: cosine {overall:.6f} after {T} steps"
  for name, c in per.items():
    assert c >= (0.999 if mode == C.GEMM_FP32_UNFUSED else 0.99), f"{name}: update cosine {c:.5f}"
  eng.close()


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("model,kw", [("vae", {}), ("scvi", {}), ("sisua", dict(n_proteins=10))])
def test_train_step_with_dropout(model, kw, mode):
  # reference defaults: input_dropout 0.3 (single_cell_model.py:80), hidden dropout 0.1 (base.yaml:10-17)
  G, B = 200, 64
  cfg, flat, mov, batch = _setup(model, kw, G, B, mode, trained_moving=False, input_dropout=0.3,
                                 enc_dropout=0.1, dec_dropout=0.1, **({"encl_dropout": 0.1} if model == "scvi" else {}))
  eng = _engine(cfg, flat, mov)
  drop = Hh.oracle_dropout_masks(cfg, B, seed=99, step=3)
  _grad_check(cfg, flat, mov, batch, eng, drop=drop, seed=99, step=3)
  # inference ignores dropout
  out = eng.infer(**batch)
  ref = O.forward(cfg, Hh.oracle_params(cfg, flat), Hh.oracle_moving(cfg, eng.bn_moving.cpu().numpy()), training=False, **batch)
  _close(out["terms"][0].cpu().numpy(), ref["elbo"].numpy(), what="elbo (inference, dropout model)")
  eng.close()


def test_errors_are_loud():
  from sisua_b200 import _lib
  cfg = C.make_step_config("scvi", n_genes=40, max_batch=64)
  from sisua_b200.engine import Engine
  eng = Engine(cfg, 0)
  b = Hh.make_batch(cfg, 16)
  with pytest.raises(_lib.SisuaError):
    eng.infer(x=b["x"], eps_z=b["eps_z"])          # scVI without library
  with pytest.raises(_lib.SisuaError):
    big = Hh.make_batch(cfg, 128)
    eng.infer(**big)                               # exceeds max_batch
  eng.close()


@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("N,K", [(96, 64), (64, 96), (80, 128)])
def test_tcgen05_descriptor_selftest(a_mn, b_mn, N, K):
  """Pins the shared-memory / instruction descriptor conventions (K-major and MN-major no-swizzle tiles, TMEM
  lane = row) that the fused kernels rely on."""
  import ctypes
  from sisua_b200 import _lib
  L = _lib.load()
  g = torch.Generator(device="cuda"); g.manual_seed(N * 1000 + K + 2 * a_mn + b_mn)
  A = torch.randn((128, K), device="cuda", generator=g)
  Bm = torch.randn((N, K), device="cuda", generator=g)
  D = torch.zeros((128, N), device="cuda")
  rc = L.sisua_tc_selftest(ctypes.c_void_p(A.data_ptr()), ctypes.c_void_p(Bm.data_ptr()), ctypes.c_void_p(D.data_ptr()),
                           N, K, a_mn, b_mn, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
  assert rc == 0
  torch.cuda.synchronize()
  ref = A.half().float() @ Bm.half().float().T
  assert torch.allclose(D, ref, rtol=1e-4, atol=1e-3), float((D - ref).abs().max())


@pytest.mark.parametrize("mode", MODES)
def test_deep_gene_pipeline(mode):
  """Many 32-gene tiles per CTA in inference and training: the chunk heuristic is overridden so that every output-head
  CTA walks all 63 gene tiles (weight / accumulator / gradient stages recycled 31 times), and the depth is asserted."""
  cfg, flat, mov, batch = _setup("vae", {}, 2000, 300, mode, trained_moving=True)
  eng = _engine(cfg, flat, mov)
  if mode == C.GEMM_TC_3XFP16:
    eng.force_chunks(out_chunks=1, enc_chunks=1, bwd_chunks=1)
    geo = eng.geometry(300)
    assert geo["out_chunks"] == 1 and geo["out_tiles_per_chunk"] == 63 and geo["enc_kblocks_per_chunk"] == 32, geo
  out = eng.infer(want_mean=True, **batch)
  torch.cuda.synchronize()
  ref = O.forward(cfg, Hh.oracle_params(cfg, flat), Hh.oracle_moving(cfg, mov), training=False, **batch)
  _close(out["terms"][0].cpu().numpy(), ref["elbo"].numpy(), what="elbo")
  _close(out["mean"].cpu().numpy(), ref["mu"].numpy(), atol=1e-7, what="imputed mean")
  _grad_check(cfg, flat, mov, batch, eng)
  eng.close()


def test_host_count_formats_roundtrip():
  """uint16 and CSR host formats unpack to exactly the fp32 matrix the step consumes."""
  from sisua_b200.engine import Engine
  from sisua_b200.pipeline import CsrBatch, quantize_counts
  cfg = C.make_step_config("vae", n_genes=558, max_batch=128)
  eng = Engine(cfg, 0)
  X = Hh.make_batch(cfg, 100)["x"]
  dst = torch.full((100, 558), -1.0, device="cuda")
  q = quantize_counts(X)
  assert q.dtype == torch.int16
  eng.unpack_counts_u16(q.cuda(), dst)
  assert torch.equal(dst.cpu(), torch.from_numpy(X))
  dst.fill_(-1.0)
  c = CsrBatch(X)
  eng.unpack_counts_csr(c.indptr.cuda(), c.cols.cuda(), c.vals.cuda(), dst)
  assert torch.equal(dst.cpu(), torch.from_numpy(X))
  assert c.nbytes < X.nbytes
  assert quantize_counts(X + 0.5).dtype == torch.float32     # non-integer data stays fp32
  # 16-bit destinations (what the step reads directly) and the packed 2-bytes-per-non-zero form
  from sisua_b200.pipeline import Csr8Batch
  X2 = X.copy()
  X2[3, :] = 0; X2[3, 557] = 7; X2[4, :] = 0; X2[5, 0] = 255; X2[5, 1] = 254; X2[5, 300] = 65535; X2[6, :] = 0; X2[6, 255] = 1; X2[6, 510] = 2
  d16 = torch.full((100, 558), -1, device="cuda", dtype=torch.int16)
  c2 = CsrBatch(X2)
  eng.unpack_counts_csr(c2.indptr.cuda(), c2.cols.cuda(), c2.vals.cuda(), d16)
  np.testing.assert_array_equal(d16.cpu().numpy().view(np.uint16).astype(np.float32), X2)
  d16.fill_(-1)
  c8 = Csr8Batch(X2)
  eng.unpack_counts_csr8(c8.indptr.cuda(), c8.big_ptr.cuda(), c8.ents.cuda(), c8.big.cuda(), d16)
  np.testing.assert_array_equal(d16.cpu().numpy().view(np.uint16).astype(np.float32), X2)
  assert c8.nbytes < 0.6 * c2.nbytes
  eng.close()


def test_cuda_graph_replay_equals_eager_steps():
  """train_step + adam_step captured as a CUDA graph and replayed == the same steps launched eagerly (dropout on:
  masks and Adam's bias correction must follow the device-side step counter, not values baked into the graph)."""
  from sisua_b200.engine import Engine
  from sisua_b200.pipeline import GraphedTrainStep
  cfg = C.make_step_config("vae", n_genes=200, max_batch=64, input_dropout=0.3, enc_dropout=0.1)
  flat = Hh.randomize_norm_params(cfg, PR.init_flat_params(cfg))
  mov = PR.init_bn_moving(cfg)
  a = Engine(cfg, 0, flat_params=flat, bn_moving=mov)
  b = Engine(cfg, 0, flat_params=flat, bn_moving=mov)
  g = GraphedTrainStep(a, 64, lr=1e-3, clipnorm=100.0, seed=5)
  losses_a, losses_b = [], []
  for t in range(1, 5):
    batch = Hh.make_batch(cfg, 64, seed=t)
    x = torch.from_numpy(batch["x"]).cuda(); e = torch.from_numpy(batch["eps_z"]).cuda()
    _, la = g.step(x, eps_z=e)
    losses_a.append(float(la))
    _, lb = b.train_step(x, eps_z=e, seed=5, step=t)
    b.adam_step(lr=1e-3, clipnorm=100.0, t=t)
    losses_b.append(float(lb))
  np.testing.assert_allclose(losses_a, losses_b, rtol=1e-6)
  # float atomics make gradient sums order-dependent at 1e-7; Adam's normalised first steps amplify that up to ~lr
  assert float((a.params - b.params).abs().max()) <= 4e-3 and float((a.params - b.params).abs().mean()) <= 1e-5
  assert losses_a[0] != losses_a[1]
  a.close(); b.close()


@pytest.mark.parametrize("model,kw,G,B", [("vae", {}, 1, 2), ("vae", {}, 7, 3), ("sisua", dict(n_proteins=1), 33, 5),
                                          ("dca", {}, 65, 129), ("scvi", {}, 31, 130)])
def test_edge_shapes(model, kw, G, B):
  """Ragged shapes: fewer genes than one tile / vector, batches that are not multiples of any tile, B = 2."""
  cfg, flat, mov, batch = _setup(model, kw, G, B, C.GEMM_TC_3XFP16, trained_moving=False)
  eng = _engine(cfg, flat, mov)
  _grad_check(cfg, flat, mov, batch, eng, gtol=8e-3)
  out = eng.infer(**batch)
  ref = O.forward(cfg, Hh.oracle_params(cfg, flat), Hh.oracle_moving(cfg, eng.bn_moving.cpu().numpy()), training=False, **batch)
  _close(out["terms"][0].cpu().numpy(), ref["elbo"].numpy(), what="elbo")
  _close(out["mean"].cpu().numpy(), ref["mu"].numpy(), atol=1e-7, what="mean")
  eng.close()


def test_wide_gene_panel_20000():
  """The 20000-gene variant of the scalability config (BASELINE.json configs[4]) at a small batch."""
  cfg, flat, mov, batch = _setup("vae", {}, 20000, 130, C.GEMM_TC_3XFP16, trained_moving=False)
  eng = _engine(cfg, flat, mov)
  terms, loss = eng.train_step(**batch)
  torch.cuda.synchronize()
  ref = O.forward(cfg, Hh.oracle_params(cfg, flat), Hh.oracle_moving(cfg, mov), training=True, **batch)
  _close(terms[0].cpu().numpy(), ref["elbo"].detach().numpy(), what="train elbo (20000 genes)")
  eng.close()


@pytest.mark.parametrize("x_dist", ["zinbd", "nbd"])
@pytest.mark.parametrize("G,B,clip", [(2100, 200, 12.0), (333, 64, 5.0)])
def test_scvi_fused_softmax_heads(x_dist, G, B, clip):
  """scVI heads on the fused kernel: gene softmax over many gene chunks (logsumexp partials), closed and open
  library clip gates (clip 5 < log library of most cells -> d loss / d library gated off), both likelihoods."""
  cfg, flat, mov, batch = _setup("scvi", dict(x_dist=x_dist, clip_library=clip), G, B, C.GEMM_TC_3XFP16, trained_moving=False)
  eng = _engine(cfg, flat, mov)
  _grad_check(cfg, flat, mov, batch, eng)
  out = eng.infer(**batch, want_disp=True)
  ref = O.forward(cfg, Hh.oracle_params(cfg, flat), Hh.oracle_moving(cfg, eng.bn_moving.cpu().numpy()), training=False, **batch)
  _close(out["terms"][0].cpu().numpy(), ref["elbo"].numpy(), what="elbo")
  _close(out["mean"].cpu().numpy(), ref["mu"].numpy(), atol=1e-7, what="mean")
  _close(out["disp"].cpu().numpy(), ref["theta"].numpy(), atol=1e-7, what="disp")
  eng.close()


def test_scvi_fused_matches_unfused_with_samples():
  """S = 3 Monte-Carlo samples through the fused scVI passes == the fp32 un-fused row kernel."""
  outs = []
  for mode in (C.GEMM_FP32_UNFUSED, C.GEMM_TC_3XFP16):
    cfg, flat, mov, _ = _setup("scvi", {}, 700, 128, mode)
    batch = Hh.make_batch(cfg, 100, S=3)
    eng = _engine(cfg, flat, mov)
    out = eng.infer(**batch, S=3, want_disp=True)
    outs.append({k: out[k].cpu().numpy() for k in ("terms", "mean", "disp")})
    eng.close()
  _close(outs[1]["terms"][0], outs[0]["terms"][0], what="elbo fused vs un-fused")
  _close(outs[1]["mean"], outs[0]["mean"], atol=1e-7, what="mean fused vs un-fused")
  _close(outs[1]["disp"], outs[0]["disp"], atol=1e-7, what="disp fused vs un-fused")


def test_scvi_reapply_activation_uses_row_kernel():
  """The literal Q2 reading (activations applied again on scVI's positive parameters) stays on the un-fused row kernel."""
  cfg, flat, mov, batch = _setup("scvi", dict(scvi_reapply_act=1), 150, 48, C.GEMM_TC_3XFP16, trained_moving=False)
  eng = _engine(cfg, flat, mov)
  _grad_check(cfg, flat, mov, batch, eng)
  eng.close()


@pytest.mark.parametrize("model,kw", [("vae", {}), ("scvi", {}), ("sisua", dict(n_proteins=10))])
@pytest.mark.parametrize("fmt", ["f32", "u16", "csr"])
def test_host_buffer_train_step_equals_device_step(model, kw, fmt):
  """sisua_train_step_host (host minibatches staged by the library, double-buffered) == sisua_train_step on the same
  batches already in HBM, over several consecutive steps (slot reuse, CSR buffers growing)."""
  from sisua_b200.engine import Engine
  from sisua_b200.pipeline import CsrBatch, quantize_counts
  cfg = C.make_step_config(model, n_genes=300, max_batch=96, input_dropout=0.2, **kw)
  flat = Hh.randomize_norm_params(cfg, PR.init_flat_params(cfg))
  mov = PR.init_bn_moving(cfg)
  a = Engine(cfg, 0, flat_params=flat, bn_moving=mov)
  b = Engine(cfg, 0, flat_params=flat, bn_moving=mov)
  host_loss = [torch.empty(1).pin_memory() for _ in range(6)]
  host_terms = [torch.empty((5, 96)).pin_memory() for _ in range(6)]
  dev_loss, dev_terms, keep = [], [], []
  for t in range(1, 7):
    Bt = 96 if t != 4 else 50                    # a ragged batch in the middle
    batch = Hh.make_batch(cfg, Bt, seed=t, stress=(t % 2 == 0))      # dense stress batches make the CSR buffers grow
    x = batch["x"]
    xh = {"f32": lambda: torch.from_numpy(x).pin_memory(), "u16": lambda: quantize_counts(x), "csr": lambda: CsrBatch(x)}[fmt]()
    extras = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in batch.items() if k != "x"}
    keep.append((xh, extras))
    a.train_step_host(xh, host_loss=host_loss[t - 1], host_terms=host_terms[t - 1][:, :Bt] if Bt == 96 else None,
                      seed=3, step=t, **extras)
    a.adam_step(lr=1e-3, clipnorm=100.0, t=t)
    tm, ls = b.train_step(**batch, seed=3, step=t)
    b.adam_step(lr=1e-3, clipnorm=100.0, t=t)
    dev_loss.append(ls.clone()); dev_terms.append(tm.clone())
  torch.cuda.synchronize()
  for t in range(6):
    # same kernels on the same data; only the order of float atomics (and, after a few Adam steps, 1e-5 parameter
    # differences) separate the two runs -- a staging bug would show up at the 1e-2 level
    np.testing.assert_allclose(float(host_loss[t]), float(dev_loss[t]), rtol=2e-4)
    if t != 3:
      np.testing.assert_allclose(host_terms[t].numpy(), dev_terms[t].cpu().numpy(), rtol=2e-3, atol=2e-3)
  assert float((a.params - b.params).abs().max()) <= 4e-3 and float((a.params - b.params).abs().mean()) <= 1e-5
  a.close(); b.close()


def test_host_buffer_step_rejects_bad_input():
  from sisua_b200 import _lib
  from sisua_b200.engine import Engine
  cfg = C.make_step_config("vae", n_genes=64, max_batch=32)
  eng = Engine(cfg, 0)
  with pytest.raises(ValueError):
    eng.train_step_host(torch.zeros((8, 63)), eps_z=torch.zeros((8, cfg.n_latent)))        # wrong gene count
  with pytest.raises(_lib.SisuaError):
    eng.train_step_host(torch.zeros((64, 64)), eps_z=torch.zeros((64, cfg.n_latent)))      # exceeds max_batch
  eng.close()


@pytest.mark.parametrize("fmt", ["f32", "u16", "csr"])
@pytest.mark.parametrize("use_graph", [True, False, "split"])
def test_host_train_pipeline_matches_device_steps(fmt, use_graph):
  """HostTrainPipeline (per-slot CUDA graphs fed by a copy stream, or the library's host-buffer entry point) trains
  exactly like explicit device-side steps on the same minibatches."""
  from sisua_b200.engine import Engine
  from sisua_b200.pipeline import CsrBatch, HostTrainPipeline, quantize_counts
  cfg = C.make_step_config("vae", n_genes=256, max_batch=64, input_dropout=0.3, enc_dropout=0.1)
  flat = Hh.randomize_norm_params(cfg, PR.init_flat_params(cfg))
  mov = PR.init_bn_moving(cfg)
  a = Engine(cfg, 0, flat_params=flat, bn_moving=mov)
  b = Engine(cfg, 0, flat_params=flat, bn_moving=mov)
  pipe = HostTrainPipeline(a, 64, use_graph=bool(use_graph))
  calls = []
  # "split": the data-parallel arrangement (two graphs with a gradient callback between them), here with one rank
  reduce_cb = (lambda g: calls.append(float(g.abs().sum()))) if use_graph == "split" else None
  la, lb, keep = [], [], []
  for t in range(1, 8):
    batch = Hh.make_batch(cfg, 64, seed=t, stress=(t == 3))
    x = batch["x"]
    xh = {"f32": lambda: torch.from_numpy(x).pin_memory(), "u16": lambda: quantize_counts(x), "csr": lambda: CsrBatch(x)}[fmt]()
    eh = torch.from_numpy(batch["eps_z"]).pin_memory()
    keep.append((xh, eh))
    out = pipe.step(xh, eh, step=t, lr=1e-3, clipnorm=100.0, allreduce=reduce_cb)
    torch.cuda.synchronize()
    la.append(float(out))
    _, ls = b.train_step(**batch, seed=0, step=t)
    b.adam_step(lr=1e-3, clipnorm=100.0, t=t)
    lb.append(float(ls))
  np.testing.assert_allclose(la, lb, rtol=2e-4)
  assert float((a.params - b.params).abs().max()) <= 4e-3 and float((a.params - b.params).abs().mean()) <= 1e-5
  assert la[0] != la[1]
  if use_graph == "split":
    assert len(calls) == 7 and all(c > 0 for c in calls)      # the callback saw this step's gradients every time
  a.close(); b.close()


@pytest.mark.parametrize("model,kw", [("vae", {}), ("scvi", {}), ("sisua", dict(n_proteins=10))])
def test_philox_reparameterisation_noise_matches_oracle(model, kw):
  """eps_z / eps_l == NULL: the kernels draw the reparameterisation noise themselves (Box-Muller on Philox4x32-10,
  regenerated by the backward pass); oracle/philox.py reproduces the same normals, so ELBO and gradients must agree."""
  from oracle.philox import NOISE_STREAM_L, NOISE_STREAM_Z, normal_noise
  G, B, seed, step = 300, 200, 0x1234567890ABCDEF, 7
  cfg, flat, mov, batch = _setup(model, kw, G, B, C.GEMM_TC_3XFP16, trained_moving=False)
  eng = _engine(cfg, flat, mov)
  inj = dict(batch)
  inj["eps_z"] = normal_noise(B, cfg.n_latent, seed, step, NOISE_STREAM_Z).astype(np.float32)
  if model == "scvi":
    inj["eps_l"] = normal_noise(B, 1, seed, step, NOISE_STREAM_L)[:, 0].astype(np.float32)
  drawn = {k: v for k, v in batch.items() if k not in ("eps_z", "eps_l")}
  terms, loss = eng.train_step(seed=seed, step=step, **drawn)
  torch.cuda.synchronize()
  z_gpu = eng.debug_buffer("z", B, cfg.n_latent).cpu().numpy()
  P = Hh.oracle_params(cfg, flat)
  for p in P.values():
    p.requires_grad_(True)
  ref = O.forward(cfg, P, Hh.oracle_moving(cfg, mov), training=True, **inj)
  ref["loss"].backward()
  _close(z_gpu, ref["z"].detach().numpy(), rtol=1e-4, atol=2e-5, what="sampled z")
  _close(terms[0].cpu().numpy(), ref["elbo"].detach().numpy(), what="elbo (Philox noise)")
  got = eng.grads_dict()
  for name, p in P.items():
    g_ref = p.grad.numpy() if p.grad is not None else np.zeros(p.shape)
    assert np.abs(got[name] - g_ref).max() <= 6e-3 * (np.abs(g_ref).max() + 1e-12) + 1e-9, name
  # a different step draws different noise; the same (seed, step) reproduces it
  t2, _ = eng.train_step(seed=seed, step=step + 1, **drawn)
  t3, _ = eng.train_step(seed=seed, step=step, **drawn)
  assert not torch.allclose(t2[0], terms[0]) and torch.allclose(t3[0], terms[0], rtol=1e-6)
  # inference: S Monte-Carlo samples, call index as the step
  S = 3
  eng.set_infer_seed(seed, 5)
  out = eng.infer(S=S, **drawn)
  inj["eps_z"] = np.stack([normal_noise(B, cfg.n_latent, seed, 5, NOISE_STREAM_Z + 2 * s) for s in range(S)]).astype(np.float32)
  if model == "scvi":
    inj["eps_l"] = np.stack([normal_noise(B, 1, seed, 5, NOISE_STREAM_L + 2 * s)[:, 0] for s in range(S)]).astype(np.float32)
  refi = O.forward(cfg, Hh.oracle_params(cfg, flat), Hh.oracle_moving(cfg, eng.bn_moving.cpu().numpy()), training=False, **inj)
  _close(out["terms"][0].cpu().numpy().reshape(S, B), refi["elbo"].numpy(), what="elbo [S,B] (Philox noise)")
  eng.close()


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("model,kw,G", [("vae", {}, 300), ("scvi", {}, 203), ("sisua", dict(n_proteins=10), 64)])
def test_row_gather_step_equals_step_on_gathered_copy(model, kw, G, mode):
  """sisua_train_step_gather (row indices into HBM-resident matrices, what fit(shuffle=True) issues) == sisua_train_step
  on an explicitly gathered copy of the same rows: ELBO terms, loss and every gradient."""
  from sisua_b200.engine import Engine
  N, B = 1000, 200
  cfg = C.make_step_config(model, n_genes=G, gemm_mode=mode, max_batch=256, input_dropout=0.3, **kw)
  flat = Hh.randomize_norm_params(cfg, PR.init_flat_params(cfg))
  mov = PR.init_bn_moving(cfg)
  full = Hh.make_batch(cfg, N, seed=4)
  rng = np.random.default_rng(0)
  rows = rng.permutation(N)[:B].astype(np.int32)
  a = Engine(cfg, 0, flat_params=flat, bn_moving=mov)
  b = Engine(cfg, 0, flat_params=flat, bn_moving=mov)
  dev = lambda v, dt=torch.float32: None if v is None else torch.from_numpy(np.ascontiguousarray(v)).to("cuda", dt)
  eps = {k: dev(full[k][:B]) for k in ("eps_z", "eps_l") if k in full}
  ta, la = a.train_step_gather(dev(full["x"]), dev(rows, torch.int32), y_all=dev(full.get("y")), library_all=dev(full.get("library")),
                               mask_all=dev(full.get("mask"), torch.uint8), seed=9, step=2, **eps)
  sub = {k: (v[rows] if k in ("x", "y", "library", "mask") else v[:B]) for k, v in full.items()}
  tb, lb = b.train_step(seed=9, step=2, **sub)
  torch.cuda.synchronize()
  np.testing.assert_allclose(ta.cpu().numpy(), tb.cpu().numpy(), rtol=1e-5, atol=1e-5)
  np.testing.assert_allclose(float(la), float(lb), rtol=1e-6)
  ga, gb = a.grads_dict(), b.grads_dict()
  # fp32 path: only the order of float atomics differs.  Fused path: the gradient GEMMs round their operands to fp16, and a
  # 1e-7 difference in an operand (atomics order upstream) can flip that rounding: fp16-grade agreement.
  tol = 1e-6 if mode == C.GEMM_FP32_UNFUSED else 3e-3
  for k in ga:
    np.testing.assert_allclose(ga[k], gb[k], rtol=1e-4 if mode == C.GEMM_FP32_UNFUSED else 0.0,
                               atol=tol * (np.abs(gb[k]).max() + 1e-12) + 1e-9, err_msg=k)
  a.close(); b.close()


@pytest.mark.parametrize("model,kw,G,mode", [("vae", {}, 2000, C.GEMM_TC_3XFP16), ("vae", dict(x_dist="nbd"), 328, C.GEMM_TC_3XFP16),
                                             ("sisua", dict(n_proteins=10), 520, C.GEMM_TC_3XFP16), ("vae", {}, 203, C.GEMM_TC_3XFP16),
                                             ("scvi", {}, 328, C.GEMM_TC_3XFP16), ("vae", {}, 328, C.GEMM_FP32_UNFUSED)])
def test_uint16_resident_counts_equal_float32(model, kw, G, mode):
  """sisua_train_step_gather_u16 (the resident shard stored as uint16, widened inside the two tcgen05 kernels) against
  sisua_train_step_gather on the float32 copy of the same matrix: same terms, loss and gradients.  G = 203 (rows not
  16-byte aligned), scVI and the un-fused mode take the staged widening instead of the in-kernel one."""
  from sisua_b200.engine import Engine
  N, B = 1000, 200
  cfg = C.make_step_config(model, n_genes=G, gemm_mode=mode, max_batch=256, input_dropout=0.3, **kw)
  flat = Hh.randomize_norm_params(cfg, PR.init_flat_params(cfg))
  mov = PR.init_bn_moving(cfg)
  full = Hh.make_batch(cfg, N, seed=4)
  full["x"][3, 5] = 4000.0; full["x"][7, G - 1] = 65535.0       # large counts survive the 16-bit storage
  rows = np.random.default_rng(0).permutation(N)[:B].astype(np.int32)
  a = Engine(cfg, 0, flat_params=flat, bn_moving=mov)
  b = Engine(cfg, 0, flat_params=flat, bn_moving=mov)
  for e in (a, b):
    e.set_count_bound(65535.0)
  dev = lambda v, dt=torch.float32: None if v is None else torch.from_numpy(np.ascontiguousarray(v)).to("cuda", dt)
  x16 = torch.from_numpy(full["x"].astype(np.uint16).view(np.int16)).cuda()
  side = dict(y_all=dev(full.get("y")), library_all=dev(full.get("library")), mask_all=dev(full.get("mask"), torch.uint8))
  ta, la = a.train_step_gather(x16, dev(rows, torch.int32), seed=9, step=2, **side)
  tb, lb = b.train_step_gather(dev(full["x"]), dev(rows, torch.int32), seed=9, step=2, **side)
  torch.cuda.synchronize()
  np.testing.assert_allclose(ta.cpu().numpy(), tb.cpu().numpy(), rtol=1e-5, atol=1e-5)
  np.testing.assert_allclose(float(la), float(lb), rtol=1e-6)
  assert np.isfinite(float(la)) and torch.isfinite(ta).all()
  ga, gb = a.grads_dict(), b.grads_dict()
  tol = 1e-6 if mode == C.GEMM_FP32_UNFUSED else 3e-3      # (atomics order + fp16 operand rounding, as in the gather test)
  for k in ga:
    np.testing.assert_allclose(ga[k], gb[k], rtol=1e-4 if mode == C.GEMM_FP32_UNFUSED else 0.0,
                               atol=tol * (np.abs(gb[k]).max() + 1e-12) + 1e-9, err_msg=k)
  np.testing.assert_array_equal(a.widen_rows(x16, dev(rows, torch.int32)).cpu().numpy(), full["x"][rows])
  np.testing.assert_array_equal(a.widen_rows(x16).cpu().numpy(), full["x"])
  a.close(); b.close()


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("model,kw", [("vae", {}), ("sisua", dict(n_proteins=10)), ("dca", {})])
@pytest.mark.parametrize("x_dist", ["zinb", "nb"])
def test_tfp_parameterised_output_enums(x_dist, model, kw, mode):
  """'zinb' / 'nb' outputs (TFP NegativeBinomial(total_count = exp(.), logits = .), tests/test_singlecell_models.py:60-80)
  through the same fused kernel as 'zinbd' / 'nbd': ELBO, parameters and every gradient vs the oracle."""
  cfg, flat, mov, batch = _setup(model, dict(kw, x_dist=x_dist), 300, 128, mode, trained_moving=False)
  # keep exp(a + b) in a sane range: shrink the random output-head weights
  d = PR.flat_to_dict(cfg, flat)
  d["out.W"] *= 0.3
  eng = _engine(cfg, flat, mov)
  _grad_check(cfg, flat, mov, batch, eng)
  out = eng.infer(**batch, want_disp=True, want_pi=True)
  ref = O.forward(cfg, Hh.oracle_params(cfg, flat), Hh.oracle_moving(cfg, eng.bn_moving.cpu().numpy()), training=False, **batch)
  _close(out["terms"][0].cpu().numpy(), ref["elbo"].numpy(), what="elbo")
  _close(out["mean"].cpu().numpy(), ref["mu"].numpy(), atol=1e-7, what="mean = total_count * exp(logits)")
  _close(out["disp"].cpu().numpy(), ref["theta"].numpy(), atol=1e-7, what="total_count")
  assert (out["pi_logit"] is not None) == (x_dist == "zinb")
  eng.close()


def _flat_device_buffers(total):
  z = lambda dt, n: torch.zeros(n, dtype=dt, device="cuda")
  return z(torch.float32, total), z(torch.float32, total), z(torch.float64, 8 * 48), z(torch.int32, 64)


@pytest.mark.parametrize("world", [1, 2, 3])
def test_peer_memory_optimizer_step_equals_allreduce_plus_adam(world):
  """sisua_adam_step_dp (reduce-scatter over peer loads -> clipnorm -> sharded Adam -> all-gather by peer stores, one
  kernel per rank, device-side barriers) against all-reduce(mean) + sisua_adam_step.  The `world` ranks are simulated on
  ONE GPU: one engine and one stream per rank, plain device buffers standing in for the symmetric memory, small grids so
  that all kernels are resident together (they wait for each other)."""
  from sisua_b200.engine import Engine
  cfg = C.make_step_config("vae", n_genes=203, max_batch=128, input_dropout=0.2)
  flat = Hh.randomize_norm_params(cfg, PR.init_flat_params(cfg)); mov = PR.init_bn_moving(cfg)
  ranks = [Engine(cfg, 0, flat_params=flat, bn_moving=mov) for _ in range(world)]
  ref = Engine(cfg, 0, flat_params=flat, bn_moving=mov)
  total = ref.total
  bufs = [_flat_device_buffers(total) for _ in range(world)]
  for r, e in enumerate(ranks):
    e.rebind(bufs[r][1], bufs[r][0])
  for r, e in enumerate(ranks):
    e.dp_bind(r, world, [b[0].data_ptr() for b in bufs], [b[1].data_ptr() for b in bufs], [b[2].data_ptr() for b in bufs],
              [b[3].data_ptr() for b in bufs], grid=24)
  streams = [torch.cuda.Stream() for _ in range(world)]
  T = 4
  for t in range(1, T + 1):
    gsum = torch.zeros(total, device="cuda")
    for r, e in enumerate(ranks):
      batch = Hh.make_batch(cfg, 96, seed=10 * t + r)
      e.train_step(seed=r, step=t, **batch)
      gsum += e.grads
    # clipnorm small enough to bite on some variables
    torch.cuda.synchronize()
    ref.grads.copy_(gsum / world)
    ref.adam_step(lr=1e-3, clipnorm=0.05, t=t)
    for r, e in enumerate(ranks):
      with torch.cuda.stream(streams[r]):
        e.adam_step_dp(lr=1e-3, clipnorm=0.05, t=t)
    torch.cuda.synchronize()
    for r, e in enumerate(ranks):
      d = (e.params - ref.params).abs().max().item()
      assert d <= 2e-6, f"step {t} rank {r}: parameters differ from all-reduce + Adam by {d:.3e}"
  # the shards tile the buffer and each rank's moments equal the reference's on its shard
  cover = torch.zeros(total, device="cuda")
  for e in ranks:
    b, en = e.dp_shard()
    cover[b:en] += 1
    assert torch.allclose(e.adam_m[b:en], ref.adam_m[b:en], rtol=1e-5, atol=1e-9)
    assert torch.allclose(e.adam_v[b:en], ref.adam_v[b:en], rtol=1e-5, atol=1e-12)
  assert bool((cover == 1).all())
  for e in ranks + [ref]:
    e.close()
