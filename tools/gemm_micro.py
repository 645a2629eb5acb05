"""micro-benchmark of single GEMM / conv launches through the C ABI (CUDA events, L2-cold inputs >> L2)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "diffusion-classifier_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
from dcb200 import engine as E
dev = torch.device("cuda:0")
ctx = E.Ctx(device=dev, precision="bf16")

import threading, time
import pynvml
pynvml.nvmlInit()
_h = pynvml.nvmlDeviceGetHandleByIndex(0)

def bench(fn, flops, n=int(os.environ.get("MICRO_ITERS", "300"))):
    """n back-to-back launches (>= 100 ms so the power governor settles); SM clock / power sampled meanwhile."""
    for _ in range(3): fn()
    torch.cuda.synchronize()
    samples, stop = [], [False]
    def pump():
        while not stop[0]:
            samples.append((pynvml.nvmlDeviceGetClockInfo(_h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(_h) / 1e3))
            time.sleep(0.005)
    th = threading.Thread(target=pump, daemon=True); th.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    stop[0] = True; th.join()
    ms = e0.elapsed_time(e1) / n
    half = samples[len(samples) // 2:] or [(0, 0)]
    clk = sorted(c for c, _ in half)[len(half) // 2]
    pw = sorted(w for _, w in half)[len(half) // 2]
    return ms, flops / ms / 1e9, clk, pw

cases = sys.argv[1:] or ["conv128", "conv256to128", "lin128", "lin256", "lin64"]
NB, H, W = 100, 128, 128
for c in cases:
    if c.startswith("conv"):
        Ci, Co = {"conv128": (128, 128), "conv256to128": (256, 128), "conv128to256": (128, 256), "conv64": (64, 64)}[c]
        x = torch.randn(NB, H, W, Ci, device=dev).to(torch.bfloat16)
        w = (torch.randn(Co, 9 * Ci, device=dev) * 0.05).to(torch.bfloat16)
        fn = lambda: E.gemm(ctx, E.conv3x3_segs(x, Ci, H, W), w, Co, NB, H, W)
        fl = 2.0 * NB * H * W * Co * 9 * Ci
    elif c.startswith("mkn"):       # mkn<M>x<K>x<N>[:act]  generic linear layer
        M, K, N = (int(v) for v in c[3:].split(":")[0].split("x"))
        x = torch.randn(M, K, device=dev).to(torch.bfloat16)
        w = (torch.randn(N, K, device=dev) * 0.05).to(torch.bfloat16)
        b = torch.randn(N, device=dev)
        act = 2 if c.endswith(":gelu") else (3 if c.endswith(":geglu") else 0)
        fn = lambda: E.linear(ctx, x, w, N, bias=b, act=act)
        fl = 2.0 * M * N * K
    else:
        N = int(c[3:])
        K = 1152
        M = NB * H * W
        x = torch.randn(M, K, device=dev).to(torch.bfloat16)
        w = (torch.randn(N, K, device=dev) * 0.05).to(torch.bfloat16)
        fn = lambda: E.linear(ctx, x, w, N)
        fl = 2.0 * M * N * K
    for dbg in os.environ.get("DBG_SWEEP", "0").split(","):
      os.environ["DCB_TC2_DBG"] = dbg
      ms, tf, clk, pw = bench(fn, fl)
      print(f"{c:14s} env={os.environ.get('DCB_TC2_DBG','0')} notc2={os.environ.get('DCB_NO_TC2','')} nohalo={os.environ.get('DCB_TC2_NO_HALO','')}  {ms:7.3f} ms  {tf:7.1f} TF/s  sm {clk} MHz {pw:.0f} W")
