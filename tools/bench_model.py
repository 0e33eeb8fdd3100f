#!/usr/bin/env python
"""Device-timed train / inference step of any model family (vae, scvi, dca, sisua) — a side tool for the
per-family numbers in profiles/; the contract bench is bench.py (ZINB-VAE, BASELINE.json's metric).

  python tools/bench_model.py --model scvi --batch 9472 --genes 2000 --steps 200
"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--model", default="scvi", choices=["vae", "scvi", "dca", "sisua"])
  ap.add_argument("--batch", type=int, default=9472)
  ap.add_argument("--genes", type=int, default=2000)
  ap.add_argument("--steps", type=int, default=200)
  ap.add_argument("--warmup", type=int, default=5)
  ap.add_argument("--gemm-mode", type=int, default=1)
  ap.add_argument("--x-dist", default="zinbd")
  a = ap.parse_args()
  import torch
  from sisua_b200 import config as C
  from sisua_b200.engine import Engine
  sys.path.insert(0, ROOT)
  from bench import synth_on_device
  dev = torch.device("cuda", 0)
  kw = dict(n_proteins=10) if a.model == "sisua" else {}
  cfg = C.make_step_config(a.model, n_genes=a.genes, gemm_mode=a.gemm_mode, max_batch=a.batch, x_dist=a.x_dist,
                           input_dropout=0.3, **kw)
  eng = Engine(cfg, 0, seed=8)
  B, G, Z, P = a.batch, a.genes, cfg.n_latent, cfg.n_proteins
  nb = max(4, (1 << 30) // (B * G * 4))
  X = synth_on_device(nb * B, G, dev, seed=1234)
  gen = torch.Generator(device=dev); gen.manual_seed(8)
  extra = {}
  if a.model != "dca":
    extra["eps_z"] = torch.randn((B, Z), device=dev, generator=gen)
  if a.model == "scvi":
    lc = torch.log(X.sum(1) + 1e-8)
    extra["library"] = torch.stack([lc.mean().expand(B), lc.var().expand(B)], 1).contiguous()
    extra["eps_l"] = torch.randn((B,), device=dev, generator=gen)
  if a.model == "sisua":
    extra["y"] = torch.poisson(torch.full((B, P), 20.0, device=dev))
    extra["mask"] = (torch.rand(B, device=dev) < 0.1).to(torch.uint8)
  terms, loss = torch.empty((5, B), device=dev), torch.empty((1,), device=dev)

  def train(i):
    j = i % nb
    eng.train_step(X[j * B:(j + 1) * B], terms=terms, loss=loss, seed=0, step=i + 1, **extra)
    eng.adam_step(lr=1e-3, clipnorm=100.0, t=i + 1)

  def timed(fn, n):
    for i in range(a.warmup):
      fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
      fn(a.warmup + i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

  eng.profile(True)
  ms_train = timed(train, a.steps)
  prof = eng.profile_read()
  eng.profile(False)
  inf_extra = {k: v for k, v in extra.items() if k not in ()}
  ms_inf = timed(lambda i: eng.infer(X[(i % nb) * B:(i % nb + 1) * B], want_mean=True, **inf_extra), max(10, a.steps // 4))
  print(json.dumps({"model": a.model, "x_dist": a.x_dist, "gemm_mode": a.gemm_mode, "batch": B, "genes": G,
                    "train_ms_per_step": ms_train, "train_cells_per_s": B / ms_train * 1e3,
                    "infer_ms_per_step": ms_inf, "infer_cells_per_s": B / ms_inf * 1e3,
                    "final_loss": float(loss.item()), "sections_ms": prof}))


if __name__ == "__main__":
  main()
