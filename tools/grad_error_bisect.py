import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from oracle import step_oracle as O
from sisua_b200 import config as C, params as PR
from sisua_b200.engine import Engine
from tests import helpers as Hh
G = 512
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 0
for B in (9472, 12032, 16384, 18880, 18944, 18944, 30000):
  cfg = C.make_step_config("vae", n_genes=G, n_latent=10, max_batch=B, gemm_mode=mode)
  flat = Hh.randomize_norm_params(cfg, PR.init_flat_params(cfg)); mov = PR.init_bn_moving(cfg)
  batch = Hh.make_batch(cfg, B, seed=1)
  P = Hh.oracle_params(cfg, flat)
  for p in P.values(): p.requires_grad_(True)
  ref = O.forward(cfg, P, Hh.oracle_moving(cfg, mov), training=True, **batch)
  ref["loss"].backward()
  for rep in range(3):
    eng = Engine(cfg, 0, flat_params=flat, bn_moving=mov)
    terms, loss = eng.train_step(seed=11, step=3, **batch)
    torch.cuda.synchronize()
    got = eng.grads_dict()
    row = {k: float(np.abs(got[k] - p.grad.numpy()).max() / (np.abs(p.grad.numpy()).max() + 1e-12)) for k, p in P.items()}
    bad = {k: v for k, v in row.items() if v > (5e-6 if mode == 0 else 5e-4)}
    print(f"B={B} mode={mode} pdl={'off' if os.environ.get('SISUA_NO_PDL') else 'on'} rep={rep} max={max(row.values()):.1e} bad={ {k: f'{v:.1e}' for k, v in bad.items()} }", flush=True)
    eng.close()
