"""Regenerates the golden fixtures tests/golden/*.npz.

The reference (trungnt13/sisua) cannot be imported in this environment (TensorFlow / odin-ai are not installed
and cannot be: DESIGN.md), so these vectors come from the float64 CPU oracle (oracle/step_oracle.py) on seeded
synthetic inputs; they freeze the oracle (and through it the CUDA path) against silent changes.  Should a
reference install ever be available, run its SingleCellModel on the stored inputs / weights and compare.

  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import step_oracle as O  # noqa: E402
from sisua_b200 import config as C  # noqa: E402
from sisua_b200 import params as PR  # noqa: E402
from tests import helpers as Hh  # noqa: E402

CASES = {
    "vae": dict(model="vae", n_genes=50, n_latent=6),
    "scvi": dict(model="scvi", n_genes=50, n_latent=6),
    "dca": dict(model="dca", n_genes=50, n_latent=6),
    "sisua": dict(model="sisua", n_genes=50, n_proteins=5, n_latent=6),
}
B = 33


def build(name):
  cfg = C.make_step_config(**CASES[name])
  flat = Hh.randomize_norm_params(cfg, PR.init_flat_params(cfg, seed=8))
  mov = PR.init_bn_moving(cfg)
  rng = np.random.default_rng(5)
  mov[:, 0, :] = rng.normal(0, 0.3, mov[:, 0, :].shape)
  mov[:, 1, :] = rng.uniform(0.5, 2.0, mov[:, 1, :].shape)
  batch = Hh.make_batch(cfg, B, seed=11)
  return cfg, flat, mov, batch


def main():
  out_dir = os.path.dirname(os.path.abspath(__file__))
  for name in CASES:
    cfg, flat, mov, batch = build(name)
    inf = O.forward(cfg, Hh.oracle_params(cfg, flat), Hh.oracle_moving(cfg, mov), training=False, **batch)
    P = Hh.oracle_params(cfg, flat)
    for p in P.values():
      p.requires_grad_(True)
    tr = O.forward(cfg, P, Hh.oracle_moving(cfg, mov), training=True, **batch)
    tr["loss"].backward()
    gnorm = {k: float(p.grad.norm()) if p.grad is not None else 0.0 for k, p in P.items()}
    np.savez_compressed(
        os.path.join(out_dir, f"{name}.npz"),
        flat_params=flat, bn_moving=mov, **{"in_" + k: v for k, v in batch.items()},
        infer_elbo=inf["elbo"].numpy(), infer_llk_x=inf["llk_x"].numpy(), infer_kl_z=inf["kl_z"].numpy(),
        infer_z_loc=inf["z_loc"].numpy(), infer_mean=inf["mu"].numpy(), infer_disp=inf["theta"].numpy(),
        train_elbo=tr["elbo"].detach().numpy(), train_loss=np.array(float(tr["loss"])),
        grad_names=np.array(list(gnorm.keys())), grad_norms=np.array(list(gnorm.values())))
    print(name, "loss", float(tr["loss"]))


if __name__ == "__main__":
  main()
