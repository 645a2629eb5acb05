"""torchrun --nproc-per-node N tests/multigpu_worker.py : worker of tests/test_gpu_g_multigpu.py (NCCL, one rank per GPU).

(1) 2-stage pruning, all classes: the (image x timestep)-sharded classify (one all-reduce per stage) must give error tables
    and labels BIT-IDENTICAL to one rank scoring everything (adding zeros is exact), identical on every rank.
(2) fast mode (random candidate classes) with ranks seeded DIFFERENTLY (seed + rank): the sharded call must equal the
    single-rank call made with rank 0's seed -- rank 0's timesteps, candidate classes and Philox seed are broadcast
    (classifier.sync_from_rank0); without that the ranks would fill different class columns."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "diffusion-classifier_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import dcb200  # noqa: E402
from helpers import CIFAR_UNET, base_cfg  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
sys.stdout.flush()
saved = os.dup(1)
os.dup2(2, 1)                       # NCCL's banner goes to stderr; stdout carries the one JSON line
dist.init_process_group("nccl", device_id=dev)
dist.barrier()
sys.stdout.flush()
os.dup2(saved, 1)
cfg = base_cfg(classes=10, n_stages=2, evaluation_per_stage=[3, 5], n_keep_per_stage=[4, 1], noise_d=32, image_size=32,
               dcb_max_batch=64, n_fast_classes=5)
torch.manual_seed(0)
dc = dcb200.DiffusionClassifier(dcb200.UNetCondition2D(**CIFAR_UNET), cfg)
with torch.no_grad():
    dc.encoder.weight.mul_(40.0)
dc = dc.to(dev).eval()
g = torch.Generator().manual_seed(1)
x = (torch.rand(5, 3, 32, 32, generator=g) * 2 - 1).to(dev)
text = torch.randint(0, 10, (5,), generator=g).to(dev)


def run(shard, seed, fast):
    cfg.dcb_shard = shard
    dc._eps_calls = 0
    torch.manual_seed(seed)
    labels = dc.classify(x, text if fast else None, fast=fast)
    return labels.clone(), dc.last_errors.clone()


def agree(t):
    got = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(got, t)
    return all(torch.equal(got[0], u) for u in got)


out = {"world": world}
l1, e1 = run(None, 77, False)                  # every rank alone
lN, eN = run("timestep", 77, False)            # sharded + one all-reduce per stage
out["two_stage_bit_identical_vs_single_rank"] = bool(torch.equal(l1, lN) and torch.equal(e1, eN))
out["two_stage_ranks_agree"] = agree(eN)
cfg.n_keep_per_stage = [2, 1]
f1, g1 = run(None, 77, True)                   # reference result: rank 0's seed, unsharded
fN, gN = run("timestep", 77 + rank, True)      # ranks seeded differently, sharded
out["fast_mode_bit_identical_vs_rank0_seed"] = bool(torch.equal(f1, fN) and torch.equal(g1, gN))
out["fast_mode_ranks_agree"] = agree(gN)
out["finite_columns_per_image_stage0"] = int(torch.isfinite(gN[:, :, 0]).sum(1).min())
# (the reference draws the wrong candidates WITH replacement, :675: between 2 and n_fast_classes distinct columns per image)
ok = all(v for k, v in out.items() if isinstance(v, bool)) and 2 <= out["finite_columns_per_image_stage0"] <= 5
if rank == 0:
    out["labels"] = lN.tolist()
    print(json.dumps(out))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"r02_multigpu_check_{world}.json"), "w"))
dist.destroy_process_group()
assert ok, out
