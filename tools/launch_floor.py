"""fixed cost of a tiny launch: per-kernel time of small GEMM / LayerNorm / GroupNorm launches replayed from a CUDA graph
(no host launch gaps), alone and interleaved (does alternating big-smem and small-smem kernels cost extra?)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "diffusion-classifier_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
from dcb200 import engine as E
dev = torch.device("cuda:0")
ctx = E.Ctx(device=dev, precision="bf16")
M = int(os.environ.get("M", "256"))
x = torch.randn(M, 512, device=dev).to(torch.bfloat16)
w = (torch.randn(512, 512, device=dev) * 0.05).to(torch.bfloat16)
b = torch.randn(512, device=dev)
g = torch.randn(512, device=dev)
xg = torch.randn(4 * 64, 512, device=dev).to(torch.bfloat16)


def run(name, fn, n=50):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(n):
            fn()
    gr.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        gr.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name:44s} {e0.elapsed_time(e1) * 1e3 / (10 * n):8.2f} us per iteration", flush=True)


gemm = lambda: E.linear(ctx, x, w, 512, bias=b)
ln = lambda: E.layernorm(ctx, x, g, b, 1e-5)
gn = lambda: E.groupnorm(ctx, xg, 512, None, 0, 4, 64, g, b, 1e-5, True)
run("tiny GEMM (gemm_tc, M=%d K=512 N=512)" % M, gemm)
run("tiny LayerNorm", ln)
run("tiny GroupNorm (fused small)", gn)
run("GEMM + LayerNorm", lambda: (gemm(), ln()))
run("GEMM + GroupNorm", lambda: (gemm(), gn()))
run("GEMM + GEMM", lambda: (gemm(), gemm()))
run("LayerNorm + LayerNorm", lambda: (ln(), ln()))
