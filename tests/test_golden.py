"""Golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py): the oracle must keep
reproducing them on the CPU, and the CUDA path must match them on the GPU."""
import os

import numpy as np
import pytest
import torch

from oracle import step_oracle as O
from sisua_b200 import config as C
from tests import helpers as Hh
from tests.golden.make_golden import CASES

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
  z = np.load(os.path.join(GOLD, f"{name}.npz"))
  cfg = C.make_step_config(**CASES[name])
  batch = {k[3:]: z[k] for k in z.files if k.startswith("in_")}
  return z, cfg, batch


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_golden(name):
  z, cfg, batch = _load(name)
  inf = O.forward(cfg, Hh.oracle_params(cfg, z["flat_params"]), Hh.oracle_moving(cfg, z["bn_moving"]), training=False, **batch)
  np.testing.assert_allclose(inf["elbo"].numpy(), z["infer_elbo"], rtol=1e-10)
  np.testing.assert_allclose(inf["mu"].numpy(), z["infer_mean"], rtol=1e-10)
  tr = O.forward(cfg, Hh.oracle_params(cfg, z["flat_params"]), Hh.oracle_moving(cfg, z["bn_moving"]), training=True, **batch)
  np.testing.assert_allclose(float(tr["loss"]), float(z["train_loss"]), rtol=1e-10)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [C.GEMM_FP32_UNFUSED, C.GEMM_TC_3XFP16])
@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_matches_golden(name, mode):
  from sisua_b200.engine import Engine
  z, cfg, batch = _load(name)
  cfg = cfg.clone(gemm_mode=mode, max_batch=256)
  eng = Engine(cfg, 0, flat_params=z["flat_params"], bn_moving=z["bn_moving"])
  out = eng.infer(want_mean=True, want_disp=True, **batch)
  torch.cuda.synchronize()
  np.testing.assert_allclose(out["terms"][0].cpu().numpy(), z["infer_elbo"], rtol=1e-4)
  np.testing.assert_allclose(out["z_loc"].cpu().numpy(), z["infer_z_loc"], rtol=1e-4, atol=1e-5)
  np.testing.assert_allclose(out["mean"].cpu().numpy(), z["infer_mean"], rtol=1e-4, atol=1e-7)
  np.testing.assert_allclose(out["disp"].cpu().numpy(), z["infer_disp"], rtol=1e-4, atol=1e-7)
  terms, loss = eng.train_step(**batch)
  torch.cuda.synchronize()
  np.testing.assert_allclose(terms[0].cpu().numpy(), z["train_elbo"], rtol=1e-4)
  np.testing.assert_allclose(float(loss), float(z["train_loss"]), rtol=1e-5)
  got = eng.grads_dict()
  gtol = 2e-3 if mode == C.GEMM_FP32_UNFUSED else 1e-2
  for n, ref in zip(z["grad_names"], z["grad_norms"]):
    g = float(np.linalg.norm(got[str(n)]))
    assert abs(g - ref) <= gtol * max(ref, 1e-6), (n, g, ref)
  eng.close()
