"""Per-GPU H2D bandwidth with all ranks copying at once (what bounds the e2e arm at N GPUs).  torchrun, GPU box."""
import os, time, torch, torch.distributed as dist
rank = int(os.environ.get("RANK", 0)); lr = int(os.environ.get("LOCAL_RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(lr)
if world > 1:
  dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
nbytes = 21 * 1024 * 1024
src = torch.empty(nbytes, dtype=torch.uint8).pin_memory(); dst = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
for _ in range(5): dst.copy_(src, non_blocking=True)
torch.cuda.synchronize()
if world > 1: dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(200): dst.copy_(src, non_blocking=True)
e1.record(); torch.cuda.synchronize()
gbs = 200 * nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9
out = torch.tensor([gbs], device="cuda")
if world > 1:
  g = [torch.zeros(1, device="cuda") for _ in range(world)]; dist.all_gather(g, out)
  if rank == 0: print("H2D GB/s per rank, all copying at once:", [round(float(x), 1) for x in g])
else:
  print("H2D GB/s alone:", round(gbs, 1))
