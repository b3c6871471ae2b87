"""torchrun entry: wall-clock sections of the sharded step (each followed by a device sync)."""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import doubly_contrastive_semseg_b200 as pkg
from doubly_contrastive_semseg_b200 import loss as L
from doubly_contrastive_semseg_b200.synthetic import WORKLOADS, make_inputs
wl = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg4"]
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
d = make_inputs(wl, seed=1, device=dev)
bl = wl.B // world; sl = slice(rank * bl, (rank + 1) * bl)
feats = d["feats"][sl].contiguous().requires_grad_(True); labels = d["labels"][sl].contiguous(); predict = d["predict"][sl].contiguous()
crit = pkg.ShardedPixelContrastLoss(device=dev); crit.max_samples, crit.max_views = wl.max_samples, wl.max_views
pc = time.perf_counter
def sync(): torch.cuda.synchronize()
acc = {}
def add(k, v): acc[k] = acc.get(k, 0.0) + v
for it in range(12):
    torch.manual_seed(it); feats.grad = None; sync(); dist.barrier(); sync()
    t0 = pc()
    B, C, h, w = feats.shape
    code, chunk, counts = L.classify(labels, predict, h, w)
    counts_all = torch.empty((world * B, 512), dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(counts_all, counts)
    ch = counts_all.cpu().numpy().reshape(world * B, 256, 2); t1 = pc()
    out = L.shard_plan_c(ch, rank, world, B, 255, wl.max_samples, wl.max_views); t2 = pc()
    sp, y_all_np = out
    lay = sp.layout
    packed = torch.from_numpy(np.concatenate([lay.req.reshape(-1), y_all_np])).to(dev)
    req_dev = packed[: lay.n_pad * 4]; y_all = packed[lay.n_pad * 4:]
    pix = L.select_pixels(code, chunk, B, h * w, req_dev, lay.n_pad); sync(); t3 = pc()
    loss = L._ShardedPixelContrastFn.apply(feats, pix, y_all, sp.n_global, 0.07, 0.07, None); sync(); t4 = pc()
    loss.backward(); sync(); t5 = pc()
    if it >= 2:
        for k, v in (("classify+gather counts+d2h", t1 - t0), ("C plan", t2 - t1), ("h2d+select", t3 - t2), ("fn.forward (sync)", t4 - t3), ("backward (sync)", t5 - t4)):
            add(k, v)
if rank == 0:
    for k, v in acc.items(): print(f"{k:30s} {v / 10 * 1e6:9.1f} us")
    print("total", sum(acc.values()) / 10 * 1e6)
dist.barrier(); dist.destroy_process_group()
