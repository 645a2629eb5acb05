"""Per-shape timing of the GEMM launches of one classify pass (CUDA events around each launch)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "diffusion-classifier_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import bench, dcb200
from dcb200 import engine as E

wl = sys.argv[1] if len(sys.argv) > 1 else "unet128"
images = int(sys.argv[2]) if len(sys.argv) > 2 else 1
arch, cfg, classes, T, gflop, ipg = bench.build_workload(wl)
if len(sys.argv) > 3:
    cfg.dcb_max_batch = int(sys.argv[3])
dev = torch.device("cuda:0")
torch.manual_seed(0)
dc = dcb200.DiffusionClassifier((dcb200.DiT if wl == "dit" else dcb200.UNetCondition2D)(**arch), cfg).to(dev).eval()
S, C = arch["sample_size"], arch["in_channels"]
x = (torch.rand(images, C, S, S) * 2 - 1).to(dev)
for _ in range(2):
    dc.classify(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); dc.classify(x); e1.record(); torch.cuda.synchronize()
print("step ms", e0.elapsed_time(e1))
cfg.dcb_cuda_graph = False
E.PROFILE = prof = E.GemmProfile()
dc.classify(x)
rows = prof.by_shape()
E.PROFILE = None
if os.environ.get("PER_LAUNCH"):
    for i, (a, b, f, tag) in enumerate(prof.rows[:int(os.environ["PER_LAUNCH"])]):
        ms = a.elapsed_time(b)
        print(f"{i:3d} {tag:34s} ms={ms:7.3f} TF/s={f/ms/1e9:8.1f}")
tot = sum(r[1] for r in rows.values())
print(f"gemm total ms {tot:.2f}")
for tag, (n, ms, fl) in sorted(rows.items(), key=lambda kv: -kv[1][1]):
    print(f"{tag:34s} n={n:3d} ms={ms:8.3f} share={ms/tot:.3f} TF/s={fl/ms/1e9:8.1f}")
