"""micro-benchmark of the fused GroupNorm + conv launches (gemm_tc2x) against the unfused pair (gn finalize + gn_apply + conv)
and the conv alone on a pre-normalised tensor; CUDA events, inputs >> L2.  DCB_TX_DBG (probes build) sweeps via DBG_SWEEP."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "diffusion-classifier_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
from dcb200 import engine as E
dev = torch.device("cuda:0")
ctx = E.Ctx(device=dev, precision="bf16")
E.FUSE_GN_MIN_C, E.FUSE_GN_MSE = 0, True     # time the fused kernel on every layer shape, whatever the product policy


def bench(fn, n=int(os.environ.get("MICRO_ITERS", "100"))):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def stats(x):
    f = x.float().reshape(-1, 128, x.shape[-1])
    return torch.stack([f.sum(1), (f * f).sum(1)], -1).contiguous()


CASES = {  # name: NB, H, W, C0, C1, div1, N, extra
    "c128": (200, 128, 128, 128, 0, 1, 128, None),
    "c128res": (200, 128, 128, 128, 0, 1, 128, "residual"),
    "c256cat": (200, 128, 128, 128, 128, 2, 128, None),
    "c128sc": (200, 128, 128, 128, 0, 1, 128, "shortcut"),
    "out3": (200, 128, 128, 128, 0, 1, 3, "mse"),
}
for name in (sys.argv[1:] or list(CASES)):
    NB, H, W, C0, C1, div1, N, extra = CASES[name]
    torch.manual_seed(0)
    x0 = torch.randn(NB * H * W, C0, device=dev).to(torch.bfloat16)
    x1 = torch.randn((NB // div1) * H * W, C1, device=dev).to(torch.bfloat16) if C1 else None
    Ct = C0 + C1
    gamma, beta = torch.randn(Ct, device=dev), torch.randn(Ct, device=dev)
    Kx = 9 * Ct
    kw, extra_segs = dict(bias=torch.randn(N, device=dev)), ()
    if extra == "residual":
        kw.update(residual=torch.randn(NB * H * W, N, device=dev).to(torch.bfloat16), res_ld=N)
    if extra == "shortcut":
        xs0 = torch.randn(NB * H * W, 128, device=dev).to(torch.bfloat16)
        xs1 = torch.randn((NB // 2) * H * W, 128, device=dev).to(torch.bfloat16)
        extra_segs = [E.seg(xs0, 128, H, W), E.seg(xs1, 128, H, W, nb_div=2)]
        Kx += 256
    w = (torch.randn(N, Kx, device=dev) * 0.02).to(torch.bfloat16)
    st0, st1 = stats(x0), (stats(x1) if C1 else None)
    if extra == "mse":
        tgt = torch.randn(NB * H * W, N, device=dev)
        err = torch.empty(NB, device=dev)
        kw.update(mse=dict(target=tgt, div=1, ld=N, err=err), want_out=False)
    else:
        kw.update(gn_stats=True)
    fl = 2.0 * NB * H * W * N * Kx

    def run(fuse):
        E.FUSE_GN = fuse
        return E.gn_conv3x3(ctx, x0, C0, x1, C1, NB, H, W, gamma, beta, 1e-5, True, w, N, div1=div1, st0=st0, st1=st1,
                            extra_segs=extra_segs, **kw)

    a = torch.randn(NB * H * W, Ct, device=dev).to(torch.bfloat16)
    conv_only = lambda: E.gemm(ctx, E.conv3x3_segs(a, Ct, H, W) + list(extra_segs), w, N, NB, H, W, **kw)
    for dbg in os.environ.get("DBG_SWEEP", "0").split(","):
        os.environ["DCB_TX_DBG"] = dbg.split(":")[0]
        os.environ.pop("DCB_TX_NBOX", None)
        if ":" in dbg:
            os.environ["DCB_TX_NBOX"] = dbg.split(":")[1]
        t_f = bench(lambda: run(True))
        print(f"{name:8s} dbg={dbg} fused {t_f:7.3f} ms {fl / t_f / 1e9:7.1f} TF/s", flush=True)
    t_u = bench(lambda: run(False))
    t_c = bench(conv_only)
    print(f"{name:8s} unfused (finalize + gn_apply + conv) {t_u:7.3f} ms {fl / t_u / 1e9:7.1f} TF/s | conv alone {t_c:7.3f} ms "
          f"{fl / t_c / 1e9:7.1f} TF/s", flush=True)
