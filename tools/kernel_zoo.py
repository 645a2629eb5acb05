"""One launch of every kernel of libdcb200.so at a shape it really runs at (unet-128 / DiT-B/4 / IPMSA bench workloads), for
`ncu --set full --profile-from-start off` (the second, profiled round sits between cudaProfilerStart/Stop).  Prints one JSON
line per launch group with the algorithmic work (bytes for the HBM-bound kernels, FLOPs for the contractions) that
tools/summarize_zoo.py divides by the measured durations.

    python tools/kernel_zoo.py > gpurun_out/zoo_plain.json &&
    ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:dcb -o gpurun_out/r02_zoo \
        python tools/kernel_zoo.py > gpurun_out/zoo.json
"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "diffusion-classifier_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import dcb200
from dcb200 import _lib as L
from dcb200 import engine as E

dev = torch.device("cuda:0")
ctx = E.Ctx(device=dev, precision="bf16")
f32 = E.Ctx(device=dev, precision="fp32")
torch.manual_seed(0)
bf = lambda *s: torch.randn(*s, device=dev).to(torch.bfloat16)
NB = 200


def stats(x):
    f = x.float().reshape(-1, 128, x.shape[-1])
    return torch.stack([f.sum(1), (f * f).sum(1)], -1).contiguous()


cases = []   # (tag, kernels expected, unit, algorithmic work, fn)


def case(tag, unit, work, fn):
    cases.append((tag, unit, work, fn))


# ---- prologue / elementwise -------------------------------------------------------------------------------------------------
x_img = torch.rand(4, 3, 128, 128, device=dev) * 2 - 1
al, sg = torch.rand(NB, device=dev), torch.rand(NB, device=dev)
img = (torch.arange(NB, device=dev) % 4).to(torch.int32)
case("prologue unet-128 (qsample + stage): 200 units 3x128x128, eps Philox, kpad 64", "bytes",
     NB * 128 * 128 * (3 * 4 * 2 + 64 * 2), lambda: E.prologue(ctx, 0, x_img, NB, 1, 3, 128, 128, 64, seed=1, alpha=al, sigma=sg,
                                                               img=img, want_target=True))
t_lab = torch.rand(NB, device=dev)
case("timestep_embed 200 x 128", "bytes", NB * 128 * 2, lambda: E.timestep_embed(ctx, t_lab, NB, 1, 128, 0))
pix = torch.rand(8, 10, 256, 256, device=dev)
case("haar_dwt [8,10,256,256] fp32", "bytes", 2 * pix.numel() * 4, lambda: dcb200.wavelet_dec_2(pix, 0.5))
wav = torch.rand(8, 40, 128, 128, device=dev)
case("haar_idwt [8,40,128,128] fp32", "bytes", 2 * wav.numel() * 4, lambda: dcb200.wavelet_enc_2(wav, 2.0))
z = torch.randn(32, 3, 128, 128, device=dev)
pred = torch.randn(64 * 128 * 128, 3, device=dev)
coef = torch.tensor([0.3, 0.8, 0.9, 0.6, 0.43, 0.2, 1.0, 0.0], device=dev)
case("ddpm_step 32 x 3x128x128, CFG pair, Philox noise", "bytes", 32 * 3 * 128 * 128 * 4 * 4,
     lambda: E.ddpm_step(f32, z, pred, 2, 0, coef, False, False, seed=3))
# ---- GroupNorm / LayerNorm ----------------------------------------------------------------------------------------------------
xg = bf(NB * 128 * 128, 128)
gam, bet = torch.randn(128, device=dev), torch.randn(128, device=dev)
st = stats(xg)
case("GroupNorm 128ch @128^2 x200 from tile statistics: gn_tiles_finalize + gn_apply", "bytes", 2 * xg.numel() * 2,
     lambda: E.groupnorm(ctx, xg, 128, None, 0, NB, 128 * 128, gam, bet, 1e-5, True, st0=st))
xs = bf(NB * 64, 1024)
g2, b2 = torch.randn(1024, device=dev), torch.randn(1024, device=dev)
case("GroupNorm 1024ch @8^2 x200: gn_stats + gn_apply", "bytes", 3 * xs.numel() * 2,
     lambda: E.groupnorm(ctx, xs, 1024, None, 0, NB, 64, g2, b2, 1e-5, True))
xc = bf(2000 * 16, 512)
g3, b3 = torch.randn(512, device=dev), torch.randn(512, device=dev)
case("GroupNorm fused small-sample mode 512ch @4^2 x2000 (CIFAR)", "bytes", 2 * xc.numel() * 2,
     lambda: E.groupnorm(ctx, xc, 512, None, 0, 2000, 16, g3, b3, 1e-5, True))
xl = bf(8 * 4096, 768)
mod = torch.randn(8, 6 * 768, device=dev)
case("LayerNorm + adaLN modulation, DiT: 8 x 4096 tokens x 768", "bytes", 2 * xl.numel() * 2,
     lambda: E.layernorm(ctx, xl, None, None, 1e-6, scale=mod[:, 768:], shift=mod, mod_ld=6 * 768, rows_per_group=4096))
xl2 = bf(NB * 256, 512)
g4, b4 = torch.randn(512, device=dev), torch.randn(512, device=dev)
case("LayerNorm affine, U-Net 16^2 level: 200 x 256 tokens x 512", "bytes", 2 * xl2.numel() * 2,
     lambda: E.layernorm(ctx, xl2, g4, b4, 1e-5))
# ---- contractions ----------------------------------------------------------------------------------------------------------------
w128 = bf(128, 9 * 128) * 0.05
case("gemm_tc2 x-halo: conv3x3 128->128 @128^2 x200 (pre-normalised input)", "flops", 2.0 * NB * 16384 * 128 * 1152,
     lambda: E.gemm(ctx, E.conv3x3_segs(xg, 128, 128, 128), w128, 128, NB, 128, 128, gn_stats=True))
x64 = bf(400 * 64 * 64, 128)
case("gemm_tc2 y-halo: conv3x3 128->128 @64^2 x400", "flops", 2.0 * 400 * 4096 * 128 * 1152,
     lambda: E.gemm(ctx, E.conv3x3_segs(x64, 128, 64, 64), w128, 128, 400, 64, 64, gn_stats=True))
xsk = bf((NB // 2) * 128 * 128, 128)
g5, b5 = torch.randn(256, device=dev), torch.randn(256, device=dev)
w256 = bf(128, 9 * 256) * 0.05
st_sk = stats(xsk)
case("gemm_tc2x (CTA pairs): GroupNorm+SiLU fused into conv3x3 cat(128,128)->128 @128^2 x200 (dominant conv) + coefficient kernel", "flops",
     2.0 * NB * 16384 * 128 * 2304,
     lambda: E.gn_conv3x3(ctx, xg, 128, xsk, 128, NB, 128, 128, g5, b5, 1e-5, True, w256, 128, div1=2, st0=st, st1=st_sk,
                          gn_stats=True))
x8 = bf(NB * 64, 1024)
w8 = bf(1024, 9 * 1024) * 0.02
case("gemm_tc: conv3x3 1024->1024 @8^2 x200 (M = 12800, K = 9216)", "flops", 2.0 * NB * 64 * 1024 * 9216,
     lambda: E.gemm(ctx, E.conv3x3_segs(x8, 1024, 8, 8), w8, 1024, NB, 8, 8))
xt = bf(NB * 256, 512)
wg = bf(4096, 512) * 0.05
bg = torch.randn(4096, device=dev)
case("gemm_tc GEGLU: [51200 x 512] -> 4096 (2048 outputs)", "flops", 2.0 * NB * 256 * 4096 * 512,
     lambda: E.linear(ctx, xt, wg, 4096, bias=bg, act=L.ACT_GEGLU))
xd = bf(32 * 4096, 768)
wq = bf(2304, 768) * 0.03
case("gemm_tc3 (CTA pairs, cta_group::2): DiT QKV [131072 x 768] -> 2304", "flops", 2.0 * 32 * 4096 * 2304 * 768,
     lambda: E.linear(ctx, xd, wq, 2304))
wf1 = bf(3072, 768) * 0.03
bf1 = torch.randn(3072, device=dev)
nws = E.attn_norms_ws(ctx, 32, 12)
case("gemm_tc3 (four epilogue groups) + attention norms in the epilogue: DiT QKV [131072 x 768] -> 2304", "flops",
     2.0 * 32 * 4096 * 2304 * 768, lambda: E.linear(ctx, xd, wq, 2304, attn_norms=(nws, 12, 4096)))
wo = bf(768, 768) * 0.03
gate_o = torch.randn(32, 768, device=dev)
case("gemm_tc3 (four epilogue groups): DiT attention out-projection, gate x out + residual [131072 x 768] -> 768", "flops",
     2.0 * 32 * 4096 * 768 * 768, lambda: E.linear(ctx, xd, wo, 768, gate=gate_o, gate_ld=768, rows_per_group=4096,
                                                   residual=xd, res_ld=768))
case("gemm_tc3 (CTA pairs, four epilogue groups): DiT FF1 + GELU [131072 x 768] -> 3072", "flops", 2.0 * 32 * 4096 * 3072 * 768,
     lambda: E.linear(ctx, xd, wf1, 3072, bias=bf1, act=L.ACT_GELU_TANH))
xo = bf(NB * 128 * 128, 128)
w3 = bf(3, 9 * 128) * 0.05
tgt = torch.randn((NB // 2) * 128 * 128, 3, device=dev)
err = torch.empty(NB, device=dev)
case("gemm_tc2 x-halo + fused eps-MSE: conv_out 128->3 @128^2 x200 + mse_finalize", "flops", 2.0 * NB * 16384 * 3 * 1152,
     lambda: E.gemm(ctx, E.conv3x3_segs(xo, 128, 128, 128), w3, 3, NB, 128, 128, want_out=False,
                    mse=dict(target=tgt, div=2, ld=3, err=err)))
# ---- attention ---------------------------------------------------------------------------------------------------------------------
qkv = bf(8 * 4096, 2304)
case("attention DiT-B/4: attn_norms + flash_attn_tc_fast (+ flash_attn_tc exits): 8 x 12 heads x 4096 x 64", "flops",
     4.0 * 8 * 12 * 4096 * 4096 * 64, lambda: E.attention(ctx, qkv, 8, 4096, 12, 64))
qkvs = qkv * 0.35     # |q| |k| inside the single-pass kernel's bound, as after DiT's LayerNorm
case("attention DiT-B/4 as the DiT blocks call it (|q||k| inside the bound, query pre-scaled): norm pass + single-pass kernel, 8 x 12 heads x 4096 x 64", "flops",
     4.0 * 8 * 12 * 4096 * 4096 * 64, lambda: E.attention(ctx, qkvs, 8, 4096, 12, 64, scale=1.0 / 1.4426950408889634))
qkv96 = bf(8 * 4096, 3 * 8 * 96)
case("attention head dim 96: flash_attn_tc_kernel<96> 8 x 8 heads x 4096 x 96", "flops", 4.0 * 8 * 8 * 4096 * 4096 * 96,
     lambda: E.attention(ctx, qkv96, 8, 4096, 8, 96))
qkv128 = bf(8 * 4096, 3 * 8 * 128)
case("attention head dim 128: flash_attn_tc_kernel<128> 8 x 8 heads x 4096 x 128", "flops", 4.0 * 8 * 8 * 4096 * 4096 * 128,
     lambda: E.attention(ctx, qkv128, 8, 4096, 8, 128))
qkv32 = bf(8 * 4096, 3 * 8 * 32)
case("attention head dim 32: flash_attn_tc_kernel<32> 8 x 8 heads x 4096 x 32", "flops", 4.0 * 8 * 8 * 4096 * 4096 * 32,
     lambda: E.attention(ctx, qkv32, 8, 4096, 8, 32))
qkv2 = bf(NB * 256, 1536)
case("attention U-Net 16^2 level: flash_attn_tc 200 x 8 heads x 256 x 64", "flops", 4.0 * NB * 8 * 256 * 256 * 64,
     lambda: E.attention(ctx, qkv2, NB, 256, 8, 64))
qkv3 = bf(NB * 64, 3072)
case("attention U-Net mid block: flash_attn_bf16 (mma.sync) 200 x 8 heads x 64 x 128", "flops", 4.0 * NB * 8 * 64 * 64 * 128,
     lambda: E.attention(ctx, qkv3, NB, 64, 8, 128))

for _, _, _, fn in cases:      # first round: attributes, tensor-map caches, allocator
    fn()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for tag, unit, work, fn in cases:
    n0 = L.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    print(json.dumps({"case": tag, "unit": unit, "work": work, "launches": L.launch_count() - n0, "event_ms": e0.elapsed_time(e1)}),
          flush=True)
torch.cuda.profiler.stop()
