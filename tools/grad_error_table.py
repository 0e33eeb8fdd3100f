"""Gradient error (max |g - g_ref| / max |g_ref| per tensor) of the CUDA train step vs the fp64 oracle for a sweep of
batch sizes and both GEMM modes (GPU box).  Usage: python tools/grad_error_table.py [model] [genes]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from oracle import step_oracle as O
from sisua_b200 import config as C, params as PR
from sisua_b200.engine import Engine
from tests import helpers as Hh

model = sys.argv[1] if len(sys.argv) > 1 else "vae"
G = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
for B in (1184, 4736, 18944):
  for mode in (0, 1):
    for drop_in in (0.0, 0.3):
      cfg = C.make_step_config(model, n_genes=G, n_latent=10, max_batch=B, input_dropout=drop_in, gemm_mode=mode)
      flat = Hh.randomize_norm_params(cfg, PR.init_flat_params(cfg)); mov = PR.init_bn_moving(cfg)
      batch = Hh.make_batch(cfg, B, seed=1)
      eng = Engine(cfg, 0, flat_params=flat, bn_moving=mov)
      terms, loss = eng.train_step(seed=11, step=3, **batch)
      torch.cuda.synchronize()
      drop = Hh.oracle_dropout_masks(cfg, B, seed=11, step=3)
      P = Hh.oracle_params(cfg, flat)
      for p in P.values(): p.requires_grad_(True)
      ref = O.forward(cfg, P, Hh.oracle_moving(cfg, mov), training=True, drop=drop, **batch)
      ref["loss"].backward()
      got = eng.grads_dict()
      row = {k: float(np.abs(got[k] - p.grad.numpy()).max() / (np.abs(p.grad.numpy()).max() + 1e-12)) for k, p in P.items()}
      dD = eng.debug_buffer("dD", B, 64).cpu().numpy()
      print(f"B={B} mode={mode} drop={drop_in} elbo_rel={float(np.abs(terms[0].cpu().numpy()-ref['elbo'].detach().numpy()).max()/np.abs(ref['elbo'].detach().numpy()).max()):.1e} " +
            " ".join(f"{k}={v:.1e}" for k, v in row.items()), flush=True)
      eng.close()
