"""torchrun --nproc-per-node N tools/multigpu_check.py : N-rank (image x timestep) sharded classify must give
bit-identical error tables and labels to the single-rank run (adding zeros in the all-reduce is exact)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "diffusion-classifier_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import torch.distributed as dist
import dcb200
from helpers import CIFAR_UNET, base_cfg

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
cfg = base_cfg(classes=10, n_stages=2, evaluation_per_stage=[3, 5], n_keep_per_stage=[4, 1], noise_d=32, image_size=32,
               dcb_max_batch=64)
torch.manual_seed(0)
dc = dcb200.DiffusionClassifier(dcb200.UNetCondition2D(**CIFAR_UNET), cfg)
with torch.no_grad():
    dc.encoder.weight.mul_(40.0)
dc = dc.to(dev).eval()
g = torch.Generator().manual_seed(1)
x = (torch.rand(5, 3, 32, 32, generator=g) * 2 - 1).to(dev)

def run(shard):
    cfg.dcb_shard = shard
    dc._eps_calls = 0
    torch.manual_seed(77)
    labels = dc.classify(x)
    return labels.clone(), dc.last_errors.clone()

l1, e1 = run(None)          # every rank alone (replica semantics)
lN, eN = run("timestep")    # sharded + one all-reduce per stage
same = bool(torch.equal(l1, lN) and torch.equal(e1, eN))
gathered = [torch.zeros_like(eN) for _ in range(world)]
dist.all_gather(gathered, eN)
ranks_agree = all(torch.equal(gathered[0], t) for t in gathered)
if rank == 0:
    out = {"world": world, "bit_identical_vs_single_rank": same, "ranks_agree": ranks_agree, "labels": lN.tolist()}
    print(json.dumps(out))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"multigpu_check_{world}.json"), "w"))
dist.destroy_process_group()
assert same and ranks_agree
