"""Host-side cost of fit(data_on='host') per step (cProfile over a short run).  GPU box."""
import cProfile, pstats, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench as BN
from sisua_b200.models import VAE, RVmeta, SingleCellData
B, G, N = 18944, 2000, 18944 * 8
X = BN.synth_on_device(N, G, torch.device("cuda", 0), seed=1).cpu().numpy()
sco = SingleCellData(X, name="bench")
m = VAE(RVmeta(G, "zinbd", True, "transcriptomic"), max_batch=B, seed=8)
m.fit(sco, batch_size=B, epochs=2, data_on="host")          # builds caches / graphs
torch.cuda.synchronize()
pr = cProfile.Profile()
timing = {"skip": 0}
pr.enable()
m.fit(sco, batch_size=B, epochs=25, data_on="host", timing=timing)
pr.disable()
print("steps", timing["steps"], "s/step", timing["seconds"] / timing["steps"])
st = pstats.Stats(pr); st.sort_stats("cumulative").print_stats(28)
