"""Timeline of out_heads_kernel's roles for CTA 0 (build with SISUA_NVCC_DEFS=-DSISUA_OUT_TRACE).  GPU box only.
Prints, per gene tile, cycles relative to the epilogue's loop top of that tile."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from sisua_b200 import config as C, _lib
from sisua_b200.engine import Engine
import bench as BN

B, G = 18944, 2000
cfg = C.make_step_config("vae", n_genes=G, n_latent=10, max_batch=B, input_dropout=0.3)
eng = Engine(cfg, 0, seed=8)
dev = torch.device("cuda", 0)
X = BN.synth_on_device(B, G, dev, seed=87654321)
terms = torch.empty((5, B), device=dev); loss = torch.empty((1,), device=dev)
for i in range(4):
  eng.train_step(X, terms=terms, loss=loss, seed=1, step=i + 1)
torch.cuda.synchronize()
lib = ctypes.CDLL(_lib.lib_path()) if hasattr(_lib, "lib_path") else _lib.load()
tr = np.zeros((4, 64, 8), dtype=np.int64)
fn = lib.sisua_debug_out_trace
fn.argtypes = [ctypes.c_void_p]; fn.restype = ctypes.c_int
assert fn(tr.ctypes.data) == 0
t0 = tr[0, 0, 0]
np.set_printoptions(linewidth=250)
print("tile | epi warp0: top  W+ACC_FULL-ok  ... (cycles since kernel's first epilogue stamp)")
names_e = ["top", "Wfull", "ACCfull", "Gfree", "math", "Gfull", "flush", "DWOfull"]
for i in range(8, 30):
  e = tr[0, i] - t0; e3 = tr[3, i] - t0; f = tr[1, i] - t0; g = tr[2, i] - t0
  print(f"tile {i:2d} EPI0 " + " ".join(f"{n}={v:7d}" for n, v in zip(names_e, e)))
  print(f"        EPI15 " + " ".join(f"{n}={v:7d}" for n, v in zip(names_e[:7], e3[:7])))
  print(f"        FWD  top={f[0]:7d} Wfull={f[1]:7d} ACCfree={f[2]:7d} issued={f[3]:7d}")
  print(f"        GRAD top={g[0]:7d} Gfull={g[1]:7d} DWOfree={g[2]:7d} issued={g[3]:7d}")
per = (tr[0, 40, 0] - tr[0, 10, 0]) / 30
print("cycles per tile (tiles 10..40):", per)
e = tr[0, 10:40]
print("EPI0 mean durations: wait W/ACC %.0f  wait Gfree %.0f  math %.0f  ->Gfull %.0f  flush %.0f (of which DWO wait %.0f)" % (
  (e[:, 2] - e[:, 0]).mean(), (e[:, 3] - e[:, 2]).mean(), (e[:, 4] - e[:, 3]).mean(), (e[:, 5] - e[:, 4]).mean(), (e[:, 6] - e[:, 5]).mean(), (e[:, 7] - e[:, 5]).mean()))
e = tr[3, 10:40]
print("EPI15 mean durations: wait W/ACC %.0f  wait Gfree %.0f  math %.0f  ->Gfull %.0f  flush %.0f" % (
  (e[:, 2] - e[:, 0]).mean(), (e[:, 3] - e[:, 2]).mean(), (e[:, 4] - e[:, 3]).mean(), (e[:, 5] - e[:, 4]).mean(), (e[:, 6] - e[:, 5]).mean()))
g = tr[2, 10:40]
print("GRAD mean: wait Gfull %.0f  wait DWOfree %.0f  issue %.0f" % ((g[:, 1] - g[:, 0]).mean(), (g[:, 2] - g[:, 1]).mean(), (g[:, 3] - g[:, 2]).mean()))
f = tr[1, 10:40]
print("FWD mean: wait Wfull %.0f  wait ACCfree %.0f  issue %.0f" % ((f[:, 1] - f[:, 0]).mean(), (f[:, 2] - f[:, 1]).mean(), (f[:, 3] - f[:, 2]).mean()))
# when did the gradient GEMMs of tile i finish?  = the moment the epilogue's G_FREE wait of tile i+2 ended, if it waited
lag = tr[0, 12:42, 3] - tr[2, 10:40, 3]
print("G_FREE observed by the epilogue of tile i+2 minus the issue of tile i's gradient MMAs: mean %.0f min %.0f max %.0f" % (lag.mean(), lag.min(), lag.max()))
