"""Ad-hoc timing of the host-format unpack kernels and of the host-buffer step's pieces."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synth_on_device
from sisua_b200 import config as C
from sisua_b200.engine import Engine
from sisua_b200.pipeline import CsrBatch, quantize_counts

B, G = 9472, 2000
dev = torch.device("cuda", 0)
cfg = C.make_step_config("vae", n_genes=G, max_batch=B, input_dropout=0.3)
eng = Engine(cfg, 0, seed=8)
X = synth_on_device(4 * B, G, dev, seed=1)
host = [X[i * B:(i + 1) * B].cpu().pin_memory() for i in range(4)]
csr = [CsrBatch(h.numpy()) for h in host]
u16 = [quantize_counts(h.numpy()) for h in host]
eps = [torch.randn(B, 10).pin_memory() for _ in range(4)]
dst = torch.empty((B, G), device=dev)
c = csr[0]
ip, cc, vv = c.indptr.cuda(), c.cols.cuda(), c.vals.cuda()
x16 = u16[0].cuda()

def timed(fn, n=50):
  for _ in range(5): fn()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(n): fn()
  e1.record(); torch.cuda.synchronize()
  return e0.elapsed_time(e1) / n * 1e3

print("unpack_csr us", timed(lambda: eng.unpack_counts_csr(ip, cc, vv, dst)), "nnz", c.cols.numel())
print("unpack_u16 us", timed(lambda: eng.unpack_counts_u16(x16, dst)))
loss = [torch.empty(1).pin_memory() for _ in range(4)]
k = [0]
def host_step(fmt):
  def f():
    i = k[0] % 4; k[0] += 1
    eng.train_step_host(fmt[i], eps_z=eps[i], host_loss=loss[i], seed=0, step=k[0])
    eng.adam_step(lr=1e-3, clipnorm=100.0, t=k[0])
  return f
def dev_step():
  i = k[0] % 4; k[0] += 1
  eng.train_step(X[i * B:(i + 1) * B], eps_z=eps_d, seed=0, step=k[0], terms=terms, loss=dl)
  eng.adam_step(lr=1e-3, clipnorm=100.0, t=k[0])
eps_d = torch.randn(B, 10, device=dev); terms = torch.empty((5, B), device=dev); dl = torch.empty(1, device=dev)
print("device step us", timed(dev_step, 200))
print("host csr step us", timed(host_step(csr), 200))
print("host u16 step us", timed(host_step(u16), 100))
t0 = time.perf_counter()
for _ in range(200): host_step(csr)()
t1 = time.perf_counter()
torch.cuda.synchronize()
print("host csr: CPU enqueue us/step", (t1 - t0) / 200 * 1e6)
def host_step_noloss(fmt):
  def f():
    i = k[0] % 4; k[0] += 1
    eng.train_step_host(fmt[i], eps_z=eps[i], host_loss=None, seed=0, step=k[0])
    eng.adam_step(lr=1e-3, clipnorm=100.0, t=k[0])
  return f
print("host csr step, no loss D2H us", timed(host_step_noloss(csr), 200))
def mixed():
  i = k[0] % 4; k[0] += 1
  eng.unpack_counts_csr(ip, cc, vv, dst)
  eng.train_step(dst, eps_z=eps_d, seed=0, step=k[0], terms=terms, loss=dl)
  eng.adam_step(lr=1e-3, clipnorm=100.0, t=k[0])
print("device csr unpack + step us", timed(mixed, 200))
def mixed2():
  i = k[0] % 4; k[0] += 1
  eng.unpack_counts_csr(ip, cc, vv, dst)
  eng.train_step(dst, eps_z=eps_d, seed=0, step=k[0], terms=terms, loss=dl)
  eng.adam_step(lr=1e-3, clipnorm=100.0, t=k[0])
  loss[i].copy_(dl, non_blocking=True)
print("device csr unpack + step + loss D2H us", timed(mixed2, 200))
side = torch.cuda.Stream()
big_host = torch.empty(10_800_000 // 4, dtype=torch.float32).pin_memory()
big_dev = torch.empty_like(big_host, device=dev)
def dev_step_with_bg_copy():
  with torch.cuda.stream(side):
    big_dev.copy_(big_host, non_blocking=True)
  dev_step()
print("device step + independent 10.8 MB H2D per step us", timed(dev_step_with_bg_copy, 200))
ev_a, ev_b = torch.cuda.Event(), torch.cuda.Event()
def dev_step_with_events():
  i = k[0] % 4; k[0] += 1
  ev_a.record()
  eng.train_step(X[i * B:(i + 1) * B], eps_z=eps_d, seed=0, step=k[0], terms=terms, loss=dl)
  ev_b.record()
  eng.adam_step(lr=1e-3, clipnorm=100.0, t=k[0])
print("device step + 2 event records us", timed(dev_step_with_events, 200))
from sisua_b200.pipeline import GraphedTrainStep
g = GraphedTrainStep(eng, B, lr=1e-3, clipnorm=100.0, seed=0)
def graph_step():
  i = k[0] % 4; k[0] += 1
  g.graph.replay()
print("graph replay step us", timed(graph_step, 200))
def graph_step_bg():
  with torch.cuda.stream(side):
    big_dev.copy_(big_host, non_blocking=True)
  g.graph.replay()
print("graph replay step + independent 10.8 MB H2D us", timed(graph_step_bg, 200))
g.x.copy_(X[:B]); g.eps_z.copy_(eps_d)
print("graph replay step (real counts) us", timed(graph_step, 200))
print("graph replay step (real counts) + independent 10.8 MB H2D us", timed(graph_step_bg, 200))
