"""torchrun entry: sharded pixel loss on G GPUs == single-GPU module on the concatenated batch.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 \
        --master-port 29541 tools/sharded_check.py [workload]
Exit code 0 iff loss (rel 1e-5) and every rank's feature gradient (rel 2e-3 of max-abs) agree."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import doubly_contrastive_semseg_b200 as pkg                                   # noqa: E402
from doubly_contrastive_semseg_b200.synthetic import WORKLOADS, make_inputs   # noqa: E402


def main():
    wl = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "small"]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    d = make_inputs(wl, seed=3, device=dev)
    bl = wl.B // world
    sl = slice(rank * bl, (rank + 1) * bl)
    crit = pkg.ShardedPixelContrastLoss(device=dev)
    crit.max_samples, crit.max_views = wl.max_samples, wl.max_views
    x = d["feats"][sl].clone().requires_grad_(True)
    torch.manual_seed(11)
    loss = crit(x, labels=d["labels"][sl].contiguous(), predict=d["predict"][sl].contiguous())
    loss.backward()
    # single-GPU module on the whole batch (every rank does it; cheap)
    ref = pkg.PixelContrastLoss(device=dev)
    ref.max_samples, ref.max_views = wl.max_samples, wl.max_views
    xf = d["feats"].clone().requires_grad_(True)
    torch.manual_seed(11)
    loss_ref = ref(xf, labels=d["labels"], predict=d["predict"])
    loss_ref.backward()
    e_loss = abs(loss.item() - loss_ref.item()) / abs(loss_ref.item())
    gref = xf.grad[sl]
    e_grad = float((x.grad - gref).abs().max() / xf.grad.abs().max())
    same_support = bool(((x.grad != 0) == (gref != 0)).all())
    ok = e_loss <= 1e-5 and e_grad <= 2e-3 and same_support and crit.last_n_global == ref.last_layout.n
    print(f"rank {rank}/{world} {wl.name}: N={crit.last_n_global} local rows {crit.last_layout.n} "
          f"loss {loss.item():.7f} vs {loss_ref.item():.7f} rel {e_loss:.1e} grad rel {e_grad:.1e} "
          f"support {same_support} -> {'OK' if ok else 'FAIL'}", flush=True)
    t = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(t)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(1 if int(t.item()) else 0)


if __name__ == "__main__":
    main()
