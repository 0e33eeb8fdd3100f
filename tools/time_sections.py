"""Per-section device time of the train step at the bench shape (GPU box): quick A/B of kernel variants.
Usage: python tools/time_sections.py [batch] [genes] [model] [dp1] [u16]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from sisua_b200 import config as C
from sisua_b200.engine import Engine
sys.path.insert(0, ROOT)
import bench as BN

B = int(sys.argv[1]) if len(sys.argv) > 1 else 18944
G = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
model = sys.argv[3] if len(sys.argv) > 3 else "vae"
kw = dict(n_proteins=10) if model == "sisua" else {}
cfg = C.make_step_config(model, n_genes=G, n_latent=10, max_batch=B, input_dropout=0.3, **kw)
eng = Engine(cfg, 0, seed=8)
dev = torch.device("cuda", 0)
X = BN.synth_on_device(4 * B, G, dev, seed=87654321)
extra = {}
if model == "scvi":
  extra["library"] = torch.tensor([[6.4, 0.08]], device=dev).repeat(B, 1)
if model == "sisua":
  extra["y"] = torch.rand((B, 10), device=dev) * 5; extra["mask"] = (torch.rand(B, device=dev) < 0.1).to(torch.uint8)
terms = torch.empty((5, B), device=dev); loss = torch.empty((1,), device=dev)
U16 = "u16" in sys.argv[4:]       # resident counts as uint16 (what bench.py / fit() use by default for count data)
Xs = X.to(torch.int32).to(torch.int16) if U16 else X
def step(i):
  eng.train_step(Xs[(i % 4) * B:(i % 4 + 1) * B], terms=terms, loss=loss, seed=1, step=i + 1, **extra)
  eng.adam_step(t=i + 1)
for i in range(5): step(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(100): step(5 + i)
e1.record(); torch.cuda.synchronize()
total = e0.elapsed_time(e1) / 100
eng.profile(True)
for i in range(50): step(200 + i)
prof = eng.profile_read()
print(f"{os.environ.get('SISUA_NVCC_DEFS', '')!r}: step {total:.4f} ms  " + "  ".join(f"{k} {v[0] / 50:.4f}" for k, v in prof.items()) + f"  loss {float(loss):.3f}", flush=True)

# ---- single-rank cost of the peer-memory optimiser kernel (world = 1: the exchange degenerates to barriers with itself)
if "dp1" in sys.argv[4:]:
  z = lambda dt, n: torch.zeros(n, dtype=dt, device=dev)
  g, p_, sq, fl = z(torch.float32, eng.total), z(torch.float32, eng.total), z(torch.float64, 8 * 48), z(torch.int32, 64)
  eng.rebind(p_, g)
  eng.dp_bind(0, 1, [g.data_ptr()], [p_.data_ptr()], [sq.data_ptr()], [fl.data_ptr()], grid=0)
  def step_dp(i):
    eng.train_step(X[(i % 4) * B:(i % 4 + 1) * B], terms=terms, loss=loss, seed=1, step=i + 1, **extra)
    eng.adam_step_dp(t=i + 1)
  for i in range(5): step_dp(1000 + i)
  eng.profile(True)
  for i in range(50): step_dp(2000 + i)
  prof = eng.profile_read()
  print("dp1 (world 1):", "  ".join(f"{k} {v[0] / 50:.4f}" for k, v in prof.items()), flush=True)
