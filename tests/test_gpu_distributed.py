"""2-rank NCCL run of the public fit() (skipped on a 1-GPU box): replicas stay identical and training works."""
import os
import socket
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import os, sys
sys.path.insert(0, os.environ["SISUA_ROOT"])
import numpy as np, torch, torch.distributed as dist
from sisua_b200 import synthetic as SY
from sisua_b200.models import VAE, RVmeta, SingleCellData
rank = int(os.environ["RANK"]); torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
d = SY.realistic_counts(1024, 96, 0, seed=5)
sco = SingleCellData(d["x"], name="toy")
m = VAE(RVmeta(96, "zinbd", True, "transcriptomic"), device=rank, max_batch=1024, seed=3)
m.fit(sco, batch_size=128, epochs=3, learning_rate=2e-3, logging_interval=1)
p = m.engine.params.clone()
ref = p.clone(); dist.broadcast(ref, src=0)
mv = m.engine.bn_moving.clone(); mref = mv.clone(); dist.broadcast(mref, src=0)
loss = np.array(m.train_history["loss"])
ok = bool(torch.equal(p, ref)) and bool(torch.allclose(mv, mref)) and np.isfinite(loss).all() and loss[-4:].mean() < loss[:4].mean()
print("RANK", rank, "OK" if ok else "FAIL", loss[0], loss[-1], flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
'''


@pytest.mark.gpu
def test_two_rank_fit_keeps_replicas_identical(tmp_path):
  if torch.cuda.device_count() < 2:
    pytest.skip("needs 2 GPUs")
  s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
  path = os.path.join(tmp_path, "run.py")
  with open(path, "w") as f:
    f.write(SCRIPT)
  env = dict(os.environ, SISUA_ROOT=ROOT)
  r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                      "127.0.0.1", "--master-port", str(port), path], env=env, capture_output=True, text=True, timeout=600)
  assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
  assert r.stdout.count("OK") == 2
