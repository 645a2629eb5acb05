"""micro-benchmark of dcb_layernorm at the DiT-B/4 adaLN shape (S x 4096 tokens x 768, per-sample scale / shift) and the
U-Net affine shape; checks the result bit for bit against the knob-free reference formula in torch fp32 (tolerance)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "diffusion-classifier_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
from dcb200 import engine as E
dev = torch.device("cuda:0")
ctx = E.Ctx(device=dev, precision="bf16")
torch.manual_seed(0)
for name, S, T, C, ada in (("DiT adaLN", 64, 4096, 768, True), ("U-Net affine", 800, 256, 512, False), ("affine 1024", 800, 64, 1024, False)):
    x = torch.randn(S * T, C, device=dev).to(torch.bfloat16)
    if ada:
        mod = torch.randn(S, 6 * C, device=dev)
        kw = dict(eps=1e-6, scale=mod[:, C:2 * C], shift=mod[:, :C], mod_ld=6 * C, rows_per_group=T)
        xf = x.float().reshape(S, T, C)
        ref = torch.nn.functional.layer_norm(xf, (C,), eps=1e-6) * (1 + mod[:, None, C:2 * C]) + mod[:, None, :C]
    else:
        g, b = torch.randn(C, device=dev), torch.randn(C, device=dev)
        kw = dict(gamma=g, beta=b, eps=1e-5)
        ref = torch.nn.functional.layer_norm(x.float(), (C,), g, b, 1e-5)
    out = E.layernorm(ctx, x, **kw)
    err = ((out.float().reshape(-1) - ref.reshape(-1)).norm() / ref.norm()).item()
    for _ in range(3): E.layernorm(ctx, x, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): E.layernorm(ctx, x, **kw)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{name}: {S*T} x {C}: {ms*1e3:.1f} us  {2*x.numel()*2/ms/1e6:.0f} GB/s  rel err {err:.2e}  sum {out.float().sum().item():.6f}")
