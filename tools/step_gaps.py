"""GPU busy time vs wall time of one graph-replayed classify step (how much of the step is host / launch gaps)."""
import os, sys, collections, re
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "diffusion-classifier_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
from torch.profiler import profile, ProfilerActivity
import bench, dcb200
wl = sys.argv[1] if len(sys.argv) > 1 else "unet128"
images = int(sys.argv[2]) if len(sys.argv) > 2 else 4
arch, cfg, classes, T, gflop, ipg = bench.build_workload(wl)
dev = torch.device("cuda:0")
torch.manual_seed(0)
dc = dcb200.DiffusionClassifier((dcb200.DiT if wl == "dit" else dcb200.UNetCondition2D)(**arch), cfg).to(dev).eval()
S, C = arch["sample_size"], arch["in_channels"]
x = (torch.rand(images, C, S, S) * 2 - 1).to(dev)
for _ in range(4):
    dc.classify(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); dc.classify(x); e1.record(); torch.cuda.synchronize()
step_ms = e0.elapsed_time(e1)
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    dc.classify(x)
    torch.cuda.synchronize()
evs = sorted([(ev.time_range.start, ev.time_range.end, ev.name) for ev in prof.events() if ev.device_type == torch.autograd.DeviceType.CUDA])
busy = sum(b - a for a, b, _ in evs) / 1e3
span = (evs[-1][1] - evs[0][0]) / 1e3
gaps = sorted([(evs[i + 1][0] - evs[i][1], evs[i][2][:40], evs[i + 1][2][:40]) for i in range(len(evs) - 1)], reverse=True)
print(f"step {step_ms:.2f} ms (events); first->last kernel span {span:.2f} ms; sum of kernel durations {busy:.2f} ms; kernels {len(evs)}")
print("largest gaps (us):")
for g, a, b in gaps[:8]:
    print(f"  {g:9.1f}  after {a}  before {b}")
print("sum of gaps > 5 us:", sum(g for g, _, _ in gaps if g > 5) / 1e3, "ms")
