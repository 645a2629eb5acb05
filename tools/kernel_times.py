"""Aggregate device time per kernel over one classify pass (torch.profiler / CUPTI)."""
import os, sys, collections, re
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "diffusion-classifier_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
from torch.profiler import profile, ProfilerActivity
import bench, dcb200

wl = sys.argv[1] if len(sys.argv) > 1 else "unet128"
images = int(sys.argv[2]) if len(sys.argv) > 2 else 1
arch, cfg, classes, T, gflop, ipg = bench.build_workload(wl)
if len(sys.argv) > 3:
    cfg.dcb_max_batch = int(sys.argv[3])
dev = torch.device("cuda:0")
torch.manual_seed(0)
dc = dcb200.DiffusionClassifier((dcb200.DiT if wl == 'dit' else dcb200.UNetCondition2D)(**arch), cfg).to(dev).eval()
S, C = arch["sample_size"], arch["in_channels"]
x = (torch.rand(images, C, S, S) * 2 - 1).to(dev)
for _ in range(2):
    dc.classify(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); dc.classify(x); e1.record(); torch.cuda.synchronize()
step_ms = e0.elapsed_time(e1)
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    dc.classify(x)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        k = re.sub(r"<.*", "", ev.name).split("(")[0]
        agg[k][0] += 1
        agg[k][1] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
tot = sum(v[1] for v in agg.values())
print(f"step {step_ms:.2f} ms; kernel time sum {tot/1e3:.2f} ms; evals {images*classes*T}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
    print(f"{k:44s} n={v[0]:4d} ms={v[1]/1e3:8.3f} share={v[1]/tot:.3f}")
