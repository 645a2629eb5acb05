"""epilogue cost probe: short-K GEMMs (the epilogue-paced ones) with / without GroupNorm tile statistics, residual,
rowvec, and with the epilogue work removed (DCB_TC2_DBG=2) -- CUDA events, inputs >> L2."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "diffusion-classifier_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
from dcb200 import engine as E
dev = torch.device("cuda:0")
ctx = E.Ctx(device=dev, precision="bf16")


def bench(fn, flops, n=40):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    return ms, flops / ms / 1e9


def case(name, M, K, N, conv=None):
    if conv:
        NB, H, W, Ci = conv
        x = torch.randn(NB, H, W, Ci, device=dev).to(torch.bfloat16)
        segs = E.conv3x3_segs(x, Ci, H, W)
        geo = (NB, H, W)
    else:
        x = torch.randn(M, K, device=dev).to(torch.bfloat16)
        segs = [E.seg(x, K, 1, M)]
        geo = (1, 1, M)
    w = (torch.randn(N, K, device=dev) * 0.05).to(torch.bfloat16)
    b = torch.randn(N, device=dev)
    res = torch.randn(M, N, device=dev).to(torch.bfloat16)
    rv = torch.randn(max(1, M // 16384), N, device=dev)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    fl = 2.0 * M * N * K
    variants = {
        "plain": dict(bias=b),
        "gn": dict(bias=b, gn_stats=True),
        "res": dict(bias=b, residual=res, res_ld=N),
        "res+gn": dict(bias=b, residual=res, res_ld=N, gn_stats=True),
        "rowvec+gn": dict(bias=b, rowvec=rv, rowvec_ld=N, rows_per_group=16384, gn_stats=True),
    }
    for vn, kw in variants.items():
        for dbg in ("0", "2"):
            os.environ["DCB_TC2_DBG"] = dbg
            ms, tf = bench(lambda: E.gemm(ctx, segs, w, N, *geo, out=out, **kw), fl)
            print(f"{name:22s} {vn:10s} dbg={dbg}  {ms:7.3f} ms  {tf:7.1f} TF/s", flush=True)
    os.environ["DCB_TC2_DBG"] = "0"


cases = sys.argv[1:] or ["lin512", "conv128", "lin768"]
for c in cases:
    if c == "lin512":
        case("lin M204800 K512 N1536", 204800, 512, 1536)
    elif c == "lin768":
        case("lin M204800 K768 N768", 204800, 768, 768)
    elif c == "conv128":
        case("conv128 100x128^2", 100 * 128 * 128, 1152, 128, conv=(100, 128, 128, 128))
    elif c == "conv64":
        case("conv128ch 400x64^2", 400 * 64 * 64, 1152, 128, conv=(400, 64, 64, 128))
