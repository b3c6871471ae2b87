"""torchrun entry: torch.profiler GPU timeline of one sharded step on rank 0."""
import os, sys
import torch, torch.distributed as dist
from torch.profiler import profile, ProfilerActivity
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import doubly_contrastive_semseg_b200 as pkg
from doubly_contrastive_semseg_b200.synthetic import WORKLOADS, make_inputs
wl = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg4"]
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
d = make_inputs(wl, seed=1, device=dev)
bl = wl.B // world; sl = slice(rank * bl, (rank + 1) * bl)
feats = d["feats"][sl].contiguous().requires_grad_(True); labels = d["labels"][sl].contiguous(); predict = d["predict"][sl].contiguous()
crit = pkg.ShardedPixelContrastLoss(device=dev); crit.max_samples, crit.max_views = wl.max_samples, wl.max_views
torch.manual_seed(1234)          # once: every rank consumes the same stream, the look-ahead stays valid
def step(s):
    feats.grad = None
    loss = crit(feats, labels=labels, predict=predict); loss.backward()
for s in range(5): step(s)
torch.cuda.synchronize(); dist.barrier()
import time
ts = []
for s in range(30):
    t0 = time.perf_counter(); step(20 + s); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e6)
if rank == 0: print("unprofiled per-step wall us (synced):", " ".join("%.0f" % v for v in ts))
dist.barrier()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for s in range(3): step(10 + s)
    torch.cuda.synchronize()
if rank == 0:
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    n = len(evs) // 3
    last = evs[-n:]
    t0 = last[0].time_range.start
    for e in last:
        print(f"{e.time_range.start - t0:9.1f} +{e.time_range.end - e.time_range.start:8.1f} us  {e.name[:80]}")
dist.barrier(); dist.destroy_process_group()
