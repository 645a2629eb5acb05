"""steady-state throughput of the bench step over a long window (the step runs against the 1 kW power cap, so 1-2 s bench
windows scatter by +-8 %): N seconds of back-to-back classify calls, evals/s per 2-second slice, SM clock / power from NVML."""
import os, sys, time, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "diffusion-classifier_b200")):
    sys.path.insert(0, p)
import torch, pynvml
import bench, dcb200
secs = float(sys.argv[1]) if len(sys.argv) > 1 else 20
wl = sys.argv[2] if len(sys.argv) > 2 else "unet128"
arch, cfg, classes, T, gflop, ipg = bench.build_workload(wl)
dev = torch.device("cuda:0")
torch.manual_seed(0)
dc = dcb200.DiffusionClassifier((dcb200.DiT if wl == "dit" else dcb200.UNetCondition2D)(**arch), cfg).to(dev).eval()
S, C = arch["sample_size"], arch["in_channels"]
x = (torch.rand(ipg, C, S, S) * 2 - 1).to(dev)
for _ in range(4):
    dc.classify(x)
torch.cuda.synchronize()
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
samples, stop = [], [False]
def pump():
    while not stop[0]:
        samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1e3))
        time.sleep(0.02)
th = threading.Thread(target=pump, daemon=True); th.start()
t0 = time.perf_counter(); slices = []; n = 0; ts = t0
while time.perf_counter() - t0 < secs:
    dc.classify(x); torch.cuda.synchronize(); n += 1
    if time.perf_counter() - ts >= 2.0:
        slices.append(round(n * ipg * classes * T / (time.perf_counter() - ts))); n = 0; ts = time.perf_counter()
stop[0] = True; th.join()
clk = sorted(c for c, _ in samples); pw = sorted(w for _, w in samples)
print(f"lib={os.environ.get('DCB_LIB','product')} {wl}: evals/s per 2 s slice {slices}; mean {sum(slices)/len(slices):.0f}; "
      f"SM clock median {clk[len(clk)//2]} MHz (p10 {clk[len(clk)//10]}, p90 {clk[9*len(clk)//10]}); power median {pw[len(pw)//2]:.0f} W (max {pw[-1]:.0f})")
