"""cProfile of the host side of one graph-replayed classify step."""
import os, sys, cProfile, pstats, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "diffusion-classifier_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import bench, dcb200
arch, cfg, classes, T, gflop, ipg = bench.build_workload("unet128")
dev = torch.device("cuda:0")
torch.manual_seed(0)
dc = dcb200.DiffusionClassifier(dcb200.UNetCondition2D(**arch), cfg).to(dev).eval()
x = (torch.rand(4, 3, 128, 128) * 2 - 1).to(dev)
for _ in range(4):
    dc.classify(x)
torch.cuda.synchronize()
t0 = time.perf_counter(); dc.classify(x); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"host time of classify() {1e3*(t1-t0):.2f} ms, until GPU done {1e3*(t2-t0):.2f} ms")
pr = cProfile.Profile(); pr.enable(); dc.classify(x); pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
