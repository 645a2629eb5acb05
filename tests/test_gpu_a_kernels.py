"""GPU parity of the HBM-bound kernels and the fp32-verify GEMM/attention engines against torch fp32 ops.
Tolerances: fp32 kernels 1e-5..1e-4 relative (north star: 1e-4 in fp32); bf16 storage 1e-2."""
import math

import pytest
import torch
import torch.nn.functional as F

from helpers import conv_ref, rel_err

pytestmark = pytest.mark.gpu


def _ctx(dev, precision):
    from dcb200 import engine as E
    return E.Ctx(device=dev, precision=precision)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_timestep_embed(dev, precision):
    from dcb200 import engine as E
    from oracle.diffusers_restated import get_timestep_embedding
    ctx = _ctx(dev, precision)
    t = torch.tensor([-14.99, -3.2, 0.0, 1.76, 15.0], device=dev)
    for dim, shift in ((128, 0.0), (256, 1.0)):
        out = E.timestep_embed(ctx, t, 5, 3, dim, shift).float()
        ref = get_timestep_embedding(t, dim, True, shift).repeat_interleave(3, 0)
        tol = 2e-5 if precision == "fp32" else 8e-3
        assert (out - ref).abs().max() < tol
    # KAT: lambda = 0 -> [1...1 | 0...0]
    out = E.timestep_embed(ctx, torch.zeros(1, device=dev), 1, 1, 128, 0.0).float()
    assert torch.equal(out[0, :64], torch.ones(64, device=dev)) and torch.equal(out[0, 64:], torch.zeros(64, device=dev))


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("C0,C1,HW,NB,silu", [(128, 0, 1024, 3, True), (256, 128, 256, 2, True), (64, 0, 64, 5, False),
                                               (1024, 1024, 64, 2, True), (256, 64, 1024, 2, True)])
def test_groupnorm(dev, precision, C0, C1, HW, NB, silu):
    from dcb200 import engine as E
    ctx = _ctx(dev, precision)
    torch.manual_seed(0)
    x0 = (torch.randn(NB, HW, C0, device=dev) * 2 + 0.5).to(ctx.tdtype)
    x1 = (torch.randn(NB, HW, C1, device=dev) - 1.0).to(ctx.tdtype) if C1 else None
    g, b = torch.randn(C0 + C1, device=dev), torch.randn(C0 + C1, device=dev)
    out = E.groupnorm(ctx, x0, C0, x1, C1, NB, HW, g, b, 1e-5, silu).float().reshape(NB, HW, C0 + C1)
    xc = x0.float() if x1 is None else torch.cat([x0.float(), x1.float()], -1)
    ref = F.group_norm(xc.permute(0, 2, 1), 32, g, b, 1e-5).permute(0, 2, 1)
    if silu:
        ref = F.silu(ref)
    assert rel_err(out, ref) < (2e-5 if precision == "fp32" else 6e-3)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("C", [256, 512, 768, 1024])
def test_layernorm(dev, precision, C):
    from dcb200 import engine as E
    ctx = _ctx(dev, precision)
    torch.manual_seed(1)
    rows, rpg = 37 * 4, 37
    x = (torch.randn(rows, C, device=dev) * 3 + 1).to(ctx.tdtype)
    g, b = torch.randn(C, device=dev), torch.randn(C, device=dev)
    tol = 2e-5 if precision == "fp32" else 6e-3
    out = E.layernorm(ctx, x, g, b, 1e-5).float()
    assert rel_err(out, F.layer_norm(x.float(), (C,), g, b, 1e-5)) < tol
    mod = torch.randn(4, 6 * C, device=dev)
    out = E.layernorm(ctx, x, None, None, 1e-6, scale=mod[:, C:], shift=mod, mod_ld=6 * C, rows_per_group=rpg).float()
    ref = F.layer_norm(x.float(), (C,), None, None, 1e-6).reshape(4, rpg, C)
    ref = (ref * (1 + mod[:, None, C:2 * C]) + mod[:, None, :C]).reshape(rows, C)
    assert rel_err(out, ref) < tol


def test_upsample_and_layout(dev):
    from dcb200 import engine as E
    for precision in ("fp32", "bf16"):
        ctx = _ctx(dev, precision)
        x = torch.randn(2, 4, 8, 64, device=dev).to(ctx.tdtype)
        up = E.upsample2x(ctx, x, 2, 4, 8, 64).reshape(2, 8, 16, 64)
        ref = F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest").permute(0, 2, 3, 1)
        assert torch.equal(up.float(), ref)
        nchw = E.nhwc_to_nchw(ctx, x.reshape(2 * 32, 64), 2, 32, 64, 64).reshape(2, 64, 4, 8)
        assert torch.equal(nchw, x.float().permute(0, 3, 1, 2))
        tok = torch.randn(2, 16, 4 * 4 * 3, device=dev).to(ctx.tdtype)
        img = E.unpatchify(ctx, tok.reshape(32, 48), 2, 4, 4, 3, 48)
        ref = torch.einsum("nhwpqc->nchpwq", tok.float().reshape(2, 4, 4, 4, 4, 3)).reshape(2, 3, 16, 16)
        assert torch.equal(img, ref)


def test_haar_matches_oracle_and_roundtrips(dev):
    import numpy as np
    from dcb200 import wavelet_dec_2, wavelet_enc_2
    from oracle import haar
    torch.manual_seed(0)
    x = torch.rand(10, 64, 48) * 2 - 1
    w = wavelet_dec_2(x.to(dev)).cpu()
    ref = haar.wavelet_dec_2_np(x.numpy())
    assert w.shape == (40, 32, 24) and np.abs(w.numpy() - ref).max() < 1e-6
    # KAT: [[1,2],[3,4]] -> cA=5, cH=-2, cV=-1, cD=0
    k = wavelet_dec_2(torch.tensor([[[1.0, 2.0], [3.0, 4.0]]], device=dev)).flatten().tolist()
    assert k == [5.0, -2.0, -1.0, 0.0]
    back = wavelet_enc_2(w.to(dev)).cpu()
    assert (back - x).abs().max() < 1e-6
    # the callers' scaling (dataset/chexpert.py:146-147: dec/2 ; plotters: enc(2*w)) at the full IPMSA size
    big = torch.rand(4, 10, 256, 256, device=dev) * 2 - 1
    assert (wavelet_enc_2(wavelet_dec_2(big, 0.5), 2.0) - big).abs().max() < 1e-6


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("v_param", [False, True])
def test_prologue_qsample(dev, mode, v_param):
    """z = alpha*x + sigma*eps (diffusion_classifier.py:115) + first-layer operand staging, pre-drawn eps."""
    from dcb200 import engine as E
    ctx = _ctx(dev, "fp32")
    torch.manual_seed(0)
    BS, C, H, W, U, rep, p = 3, 3, 8, 8, 5, 2, 2
    x = torch.rand(BS, C, H, W, device=dev) * 2 - 1
    eps = torch.randn(U, C, H, W, device=dev)
    alpha, sigma = torch.rand(U, device=dev), torch.rand(U, device=dev)
    img = torch.tensor([0, 2, 1, 1, 0], device=dev, dtype=torch.int32)
    kpad = 64
    a, tgt = E.prologue(ctx, mode, x, U, rep, C, H, W, kpad, patch=p, eps=eps, alpha=alpha, sigma=sigma, img=img,
                        want_target=True, v_param=v_param)
    z = alpha.view(-1, 1, 1, 1) * x[img.long()] + sigma.view(-1, 1, 1, 1) * eps
    t_ref = eps - sigma.view(-1, 1, 1, 1) * z if v_param else eps
    if mode == 0:
        cols = F.unfold(z, 3, padding=1).reshape(U, C, 9, H * W).permute(0, 3, 2, 1).reshape(U, H * W, 9 * C)
        t_ref = t_ref.permute(0, 2, 3, 1).reshape(-1)
    else:
        cols = F.unfold(z, p, stride=p).reshape(U, C, p * p, -1).permute(0, 3, 2, 1).reshape(U, -1, p * p * C)
        t_ref = F.unfold(t_ref, p, stride=p).reshape(U, C, p * p, -1).permute(0, 3, 2, 1).reshape(-1)
    rows = cols.shape[1]
    a = a.reshape(U, rep, rows, kpad)
    assert (a[:, 0, :, :cols.shape[2]] - cols).abs().max() < 1e-6 and torch.equal(a[:, 0], a[:, 1])
    assert a[..., cols.shape[2]:].abs().max() == 0
    assert (tgt - t_ref).abs().max() < 1e-6


def test_prologue_philox_statistics(dev):
    from dcb200 import engine as E
    ctx = _ctx(dev, "fp32")
    U, C, H, W = 8, 3, 64, 64
    x = torch.zeros(1, C, H, W, device=dev)
    one = torch.ones(U, device=dev)
    img = torch.zeros(U, device=dev, dtype=torch.int32)
    _, e1 = E.prologue(ctx, 0, x, U, 1, C, H, W, 64, seed=7, unit_id0=100, alpha=one, sigma=one, img=img, want_target=True)
    _, e2 = E.prologue(ctx, 0, x, U, 1, C, H, W, 64, seed=7, unit_id0=100, alpha=one, sigma=one, img=img, want_target=True)
    _, e3 = E.prologue(ctx, 0, x, U, 1, C, H, W, 64, seed=7, unit_id0=104, alpha=one, sigma=one, img=img, want_target=True)
    assert torch.equal(e1, e2)                               # counter-based: reproducible
    n = C * H * W
    assert torch.equal(e1[4 * n:], e3[:4 * n])               # keyed by global unit id, not by launch
    assert abs(float(e1.mean())) < 0.02 and abs(float(e1.std()) - 1) < 0.02
    assert abs(float((e1 ** 4).mean()) - 3.0) < 0.15


def test_eps_mse(dev):
    from dcb200 import engine as E
    ctx = _ctx(dev, "fp32")
    S, div, K = 6, 2, 3 * 32 * 32
    pred, tgt = torch.randn(S, K, device=dev), torch.randn(S // div, K, device=dev)
    scale = torch.rand(S, device=dev)
    err = torch.empty(S, device=dev)
    E.eps_mse(ctx, pred, tgt, scale, S, div, K, err)
    ref = (torch.norm((scale[:, None] * pred - tgt.repeat_interleave(div, 0)), dim=1, p=2) ** 2)
    assert rel_err(err, ref) < 1e-5


@pytest.mark.parametrize("case", ["conv3x3", "stride2", "concat_shortcut", "linear_epilogue", "geglu"])
def test_gemm_simt_fp32(dev, case):
    from dcb200 import _lib as L
    from dcb200 import engine as E
    ctx = _ctx(dev, "fp32")
    torch.manual_seed(0)
    if case in ("conv3x3", "stride2"):
        NB, H, W, Ci, Co = 2, 8, 16, 32, 48
        st = 1 if case == "conv3x3" else 2
        x = torch.randn(NB, H, W, Ci, device=dev)
        w, b = torch.randn(Co, Ci, 3, 3, device=dev) * 0.1, torch.randn(Co, device=dev)
        wp = w.permute(0, 2, 3, 1).reshape(Co, -1).contiguous()
        out = E.gemm(ctx, E.conv3x3_segs(x, Ci, H, W, st), wp, Co, NB, H // st, W // st, bias=b)
        assert rel_err(out.reshape(NB, H // st, W // st, Co), conv_ref(x, w, b, st)) < 1e-5
    elif case == "concat_shortcut":
        NB, H, W, C0, C1, Co = 2, 8, 8, 32, 16, 32
        a2, x0, x1 = (torch.randn(NB, H, W, c, device=dev) for c in (Co, C0, C1))
        w2, ws = torch.randn(Co, Co, 3, 3, device=dev) * 0.1, torch.randn(Co, C0 + C1, 1, 1, device=dev) * 0.1
        b = torch.randn(Co, device=dev)
        wp = torch.cat([w2.permute(0, 2, 3, 1).reshape(Co, -1), ws.reshape(Co, -1)], 1).contiguous()
        segs = E.conv3x3_segs(a2, Co, H, W) + [E.seg(x0, C0, H, W), E.seg(x1, C1, H, W)]
        out = E.gemm(ctx, segs, wp, Co, NB, H, W, bias=b)
        ref = conv_ref(a2, w2, b) + conv_ref(torch.cat([x0, x1], -1), ws, None, 1, 0)
        assert rel_err(out.reshape(NB, H, W, Co), ref) < 1e-5
    elif case == "linear_epilogue":
        M, K, N, rpg = 96, 64, 80, 32
        x, w, b = torch.randn(M, K, device=dev), torch.randn(N, K, device=dev) * 0.1, torch.randn(N, device=dev)
        rv, gate, res = torch.randn(4, N, device=dev), torch.randn(3, N, device=dev), torch.randn(M, N, device=dev)
        idx = torch.tensor([3, 0, 2], device=dev, dtype=torch.int32)
        out = E.linear(ctx, x, w, N, bias=b, rowvec=rv, rowvec_ld=N, rowvec_idx=idx, rows_per_group=rpg, gate=gate,
                       gate_ld=N, residual=res, res_ld=N, act=L.ACT_GELU_TANH, act_post=L.ACT_SILU)
        grp = torch.arange(M, device=dev) // rpg
        v = x @ w.t() + b + rv[idx.long()][grp]
        ref = F.silu(F.gelu(v, approximate="tanh") * gate[grp] + res)
        assert rel_err(out, ref) < 1e-5
    else:
        M, K, inner = 70, 64, 256
        x = torch.randn(M, K, device=dev)
        w, b = torch.randn(2 * inner, K, device=dev) * 0.1, torch.randn(2 * inner, device=dev)
        wp = torch.cat([w[:inner].reshape(-1, 128, K), w[inner:].reshape(-1, 128, K)], 1).reshape(2 * inner, K)
        bp = torch.cat([b[:inner].reshape(-1, 128), b[inner:].reshape(-1, 128)], 1).reshape(-1)
        out = E.linear(ctx, x, wp.contiguous(), 2 * inner, bias=bp.contiguous(), act=L.ACT_GEGLU)
        h, g = (x @ w.t() + b).chunk(2, -1)
        assert out.shape == (M, inner) and rel_err(out, h * F.gelu(g)) < 1e-5


@pytest.mark.parametrize("d", [32, 64, 96, 128])
@pytest.mark.parametrize("N", [16, 64, 200, 1024])
def test_attention(dev, d, N):
    """softmax(QK^T/sqrt(d))V == torch SDPA (diffusers AttnProcessor2_0); fp32 engine 1e-5, bf16 flash engine 1e-2."""
    from dcb200 import engine as E
    torch.manual_seed(0)
    B, heads = 2, 3
    qkv = torch.randn(B * N, 3 * heads * d, device=dev)
    q, k, v = (t.reshape(B, N, heads, d).transpose(1, 2) for t in qkv.chunk(3, -1))
    ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B * N, heads * d)
    out = E.attention(_ctx(dev, "fp32"), qkv, B, N, heads, d)
    assert rel_err(out, ref) < 2e-5
    qb = qkv.to(torch.bfloat16)
    q, k, v = (t.float().reshape(B, N, heads, d).transpose(1, 2) for t in qb.chunk(3, -1))
    ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B * N, heads * d)
    out = E.attention(_ctx(dev, "bf16"), qb, B, N, heads, d).float()
    assert rel_err(out, ref) < 1e-2
    out = E.attention(_ctx(dev, "bf16"), qb, B, N, heads, d, simt=True).float()
    assert rel_err(out, ref) < 1e-2


@pytest.mark.parametrize("dt", ["bf16", "fp32"])
def test_groupnorm_unit_divisor_and_expand(dev, dt):
    """GroupNorm over cat([per-sample h, per-unit skip]) with a sample divisor == GroupNorm over the materialised
    expansion; dcb_expand_samples == repeat_interleave."""
    from dcb200 import engine as E
    torch.manual_seed(0)
    ctx = E.Ctx(device=dev, precision=dt)
    U, rep, HW, C0, C1 = 3, 4, 256, 128, 64
    S = U * rep
    h = torch.randn(S * HW, C0, device=dev).to(ctx.tdtype)
    sk = torch.randn(U * HW, C1, device=dev).to(ctx.tdtype)
    g, b = torch.randn(C0 + C1, device=dev), torch.randn(C0 + C1, device=dev)
    ex = E.expand_samples(ctx, sk, S, rep, HW)
    assert torch.equal(ex, sk.reshape(U, HW, C1).repeat_interleave(rep, 0).reshape(-1, C1))
    a = E.groupnorm(ctx, h, C0, sk, C1, S, HW, g, b, 1e-5, True, div1=rep)
    r = E.groupnorm(ctx, h, C0, ex, C1, S, HW, g, b, 1e-5, True)
    assert torch.equal(a, r)
    a = E.groupnorm(ctx, sk, C1, None, 0, S, HW, g[:C1].contiguous(), b[:C1].contiguous(), 1e-6, False, div0=rep)
    r = E.groupnorm(ctx, ex, C1, None, 0, S, HW, g[:C1].contiguous(), b[:C1].contiguous(), 1e-6, False)
    assert torch.equal(a, r)


@pytest.mark.parametrize("B,heads,N,qscale", [(2, 3, 128, 3.0), (1, 2, 256, 3.0), (2, 12, 300, 3.0), (1, 8, 4096, 3.0),
                                               (3, 1, 1000, 3.0), (2, 4, 1100, 1.0), (2, 4, 1024, 12.0), (1, 2, 2048, 0.0)])
def test_attention_tcgen05_head64(dev, B, heads, N, qscale):
    """the tcgen05/TMEM flash kernel (d = 64, N >= 128: DiT-B/4 and the C/8 = 64 U-Net blocks) against torch SDPA in
    fp32 on the same bf16 inputs, against the mma.sync kernel it replaces, and for ragged N (masked last key block,
    partially empty query tiles).  N >= 1024 takes the single-pass kernel (softmax reference = |q| max|k|) unless the
    logits are so large (qscale 12) that the device falls back to the running-maximum kernel; qscale 0 = all-zero
    queries (uniform attention)."""
    import os
    from dcb200 import engine as E
    torch.manual_seed(N)
    d = 64
    qkv = torch.randn(B * N, 3 * heads * d, device=dev)
    qkv[:, : heads * d] *= qscale       # sharper softmax: the maximum moves between key blocks
    qb = qkv.to(torch.bfloat16)
    q, k, v = (t.float().reshape(B, N, heads, d).transpose(1, 2) for t in qb.chunk(3, -1))
    ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B * N, heads * d)
    out = E.attention(_ctx(dev, "bf16"), qb, B, N, heads, d).float()
    assert torch.isfinite(out).all()
    assert rel_err(out, ref) < 1e-2
    simt = E.attention(_ctx(dev, "bf16"), qb, B, N, heads, d, simt=True).float()
    assert rel_err(out, simt) < 1e-2


@pytest.mark.parametrize("B,heads,N", [(5, 12, 1024), (1, 4, 1024), (2, 12, 4096), (5, 12, 1152)])
def test_attention_norm_prepass_folded_into_the_qkv_projection(dev, B, heads, N):
    """dcb_gemm_desc.attn_norms: the projection that writes q and k leaves max |q_i|^2 / max |k_j|^2 per (sample, head) --
    from the CTA-pair kernel's epilogue registers (first and last case: >= 148 tiles) or from the row pass the library
    runs by itself for every other kernel (middle case) -- and dcb_attention_ws(NORMS_READY) gives bit for bit the result
    of the launch that computes them itself."""
    from dcb200 import engine as E
    torch.manual_seed(B * N)
    ctx = _ctx(dev, "bf16")
    D = heads * 64
    x = torch.randn(B * N, D, device=dev).to(torch.bfloat16)
    w = (torch.randn(3 * D, D, device=dev) * D ** -0.5).to(torch.bfloat16)
    b = torch.randn(3 * D, device=dev)
    ws = E.attn_norms_ws(ctx, B, heads)
    ws.fill_(7.0)                                   # the call zeroes it
    qkv = E.linear(ctx, x, w, 3 * D, bias=b, attn_norms=(ws, heads, N))
    assert torch.equal(qkv, E.linear(ctx, x, w, 3 * D, bias=b))
    qk = qkv[:, : 2 * D].float().reshape(B, N, 2, heads, 64)
    want = qk.pow(2).sum(-1).amax(1).permute(0, 2, 1)            # [B, heads, {q, k}]
    got = ws[2:].reshape(B, heads, 2)
    assert torch.allclose(got, want, rtol=1e-2), (got - want).abs().max()
    a = E.attention(ctx, qkv, B, N, heads, 64, norms_ready=ws)
    assert torch.equal(a, E.attention(ctx, qkv, B, N, heads, 64))


@pytest.mark.parametrize("d", [32, 96, 128])
@pytest.mark.parametrize("B,heads,N,qscale", [(2, 3, 128, 3.0), (2, 5, 300, 3.0), (1, 4, 1024, 3.0), (1, 2, 4096, 1.0),
                                               (3, 1, 1000, 6.0), (1, 2, 512, 0.0)])
def test_attention_tcgen05_other_head_dims(dev, d, B, heads, N, qscale):
    """the running-maximum tcgen05 kernel at head dims 32 / 96 / 128 (two 64-channel boxes per tile for d > 64, P written
    over S in TMEM, P V before the next Q K^T) against torch SDPA in fp32 on the same bf16 inputs and against the
    mma.sync kernel (DCB_KNOB_ATTN_NO_TC) it replaces for N >= 128; ragged N, sharp softmax, all-zero queries."""
    from dcb200 import engine as E, _lib as L
    torch.manual_seed(N + d)
    qkv = torch.randn(B * N, 3 * heads * d, device=dev)
    qkv[:, : heads * d] *= qscale
    qb = qkv.to(torch.bfloat16)
    q, k, v = (t.float().reshape(B, N, heads, d).transpose(1, 2) for t in qb.chunk(3, -1))
    ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B * N, heads * d)
    out = E.attention(_ctx(dev, "bf16"), qb, B, N, heads, d).float()
    assert torch.isfinite(out).all()
    assert rel_err(out, ref) < 1e-2
    with L.knob("ATTN_NO_TC"):
        old = E.attention(_ctx(dev, "bf16"), qb, B, N, heads, d).float()
    assert rel_err(out, old) < 1e-2
    again = E.attention(_ctx(dev, "bf16"), qb, B, N, heads, d).float()
    assert torch.equal(out, again)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("C0,C1,HW,NB,div1", [(256, 0, 64, 7, 1), (256, 256, 16, 300, 1), (512, 0, 16, 33, 1),
                                              (384, 128, 64, 12, 3), (1024, 0, 16, 5, 1)])
def test_groupnorm_fused_small_samples(dev, precision, C0, C1, HW, NB, div1):
    """dcb_groupnorm_fused (HW < 128: statistics + apply in one launch, one block per sample) vs torch and vs the
    two-kernel path; a sample's result does not depend on the batch it sits in."""
    from dcb200 import engine as E
    ctx = _ctx(dev, precision)
    torch.manual_seed(0)
    x0 = (torch.randn(NB, HW, C0, device=dev) * 2 + 0.5).to(ctx.tdtype)
    x1 = (torch.randn(NB // div1, HW, C1, device=dev) - 1.0).to(ctx.tdtype) if C1 else None
    g, b = torch.randn(C0 + C1, device=dev), torch.randn(C0 + C1, device=dev)
    assert E.USE_FUSED_SMALL_GN
    out = E.groupnorm(ctx, x0, C0, x1, C1, NB, HW, g, b, 1e-5, True, div1=div1)
    xc = x0.float() if x1 is None else torch.cat([x0.float(), x1.float().repeat_interleave(div1, 0)], -1)
    ref = F.silu(F.group_norm(xc.permute(0, 2, 1), 32, g, b, 1e-5).permute(0, 2, 1))
    tol = 2e-5 if precision == "fp32" else 6e-3
    assert rel_err(out.float().reshape(NB, HW, -1), ref) < tol
    E.USE_FUSED_SMALL_GN = False
    try:
        two = E.groupnorm(ctx, x0, C0, x1, C1, NB, HW, g, b, 1e-5, True, div1=div1)
    finally:
        E.USE_FUSED_SMALL_GN = True
    assert rel_err(out.float(), two.float()) < (1e-6 if precision == "fp32" else 3e-3)
    k = div1 * max(1, NB // div1 // 2)       # a prefix of the batch (whole units) gives bit-identical rows
    part = E.groupnorm(ctx, x0[:k].contiguous(), C0, None if x1 is None else x1[:k // div1].contiguous(), C1, k, HW, g, b, 1e-5,
                       True, div1=div1)
    assert torch.equal(part, out[:k * HW])


def test_attention_tcgen05_fast_path_decided_per_head(dev):
    """single-pass vs running-maximum kernel is chosen per (batch, head) from that head's own norms: a sample's attention
    output is bit-identical whether or not a sample with huge logits (which forces the fallback for ITS heads) shares the
    launch."""
    from dcb200 import engine as E
    torch.manual_seed(0)
    B, heads, N, d = 3, 4, 1024, 64
    qkv = torch.randn(B * N, 3 * heads * d, device=dev)
    qkv[N:2 * N, :heads * d] *= 14.0          # sample 1: logits far beyond the single-pass kernel's exactness bound
    qb = qkv.to(torch.bfloat16)
    q, k, v = (t.float().reshape(B, N, heads, d).transpose(1, 2) for t in qb.chunk(3, -1))
    ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B * N, heads * d)
    out = E.attention(_ctx(dev, "bf16"), qb, B, N, heads, d)
    assert torch.isfinite(out).all() and rel_err(out.float(), ref) < 1e-2
    alone0 = E.attention(_ctx(dev, "bf16"), qb[:N].contiguous(), 1, N, heads, d)
    alone2 = E.attention(_ctx(dev, "bf16"), qb[2 * N:].contiguous(), 1, N, heads, d)
    assert torch.equal(out[:N], alone0) and torch.equal(out[2 * N:], alone2)


def test_attention_fallback_kernel_strides_over_samples(dev):
    """as the fallback of the single-pass kernel the running-maximum kernel is launched with 8 CTAs per (query tile, head)
    that stride over the samples and re-initialise their barriers per sample: with 20 samples, three of which (3, 11, 19 --
    the same CTA) and one more (6) have logits beyond the single-pass bound in some heads, every sample still equals its
    own single-sample launch bit for bit and torch SDPA within tolerance."""
    from dcb200 import engine as E
    torch.manual_seed(1)
    B, heads, N, d = 20, 2, 1024, 64
    qkv = torch.randn(B * N, 3 * heads * d, device=dev)
    for smp, hd in ((3, 0), (11, 1), (19, 0), (19, 1), (6, 1)):
        qkv[smp * N:(smp + 1) * N, hd * d:(hd + 1) * d] *= 14.0
    qb = qkv.to(torch.bfloat16)
    q, k, v = (t.float().reshape(B, N, heads, d).transpose(1, 2) for t in qb.chunk(3, -1))
    ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B * N, heads * d)
    out = E.attention(_ctx(dev, "bf16"), qb, B, N, heads, d)
    assert torch.isfinite(out).all() and rel_err(out.float(), ref) < 1e-2
    for smp in (0, 3, 6, 11, 19):
        alone = E.attention(_ctx(dev, "bf16"), qb[smp * N:(smp + 1) * N].contiguous(), 1, N, heads, d)
        assert torch.equal(out[smp * N:(smp + 1) * N], alone), smp


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_pack_entry_points_match_the_torch_layout_code(dev, precision):
    """dcb_pack_conv / dcb_pack_geglu / dcb_pack_upsample / dcb_pack_rows (the C-ABI packers a non-Python binder uses) against
    the plain torch layout code they replaced: bit-identical (same fp32 sums in the same order, same rounding)."""
    from dcb200 import engine as E
    torch.manual_seed(0)
    ctx = E.Ctx(device=dev, precision=precision)
    dt = ctx.tdtype
    w = torch.randn(48, 20, 3, 3, device=dev)
    ref = torch.zeros(48, 192, device=dev)
    ref[:, :180] = w.permute(0, 2, 3, 1).reshape(48, -1)
    assert torch.equal(E.pack_conv(ctx, w, 192), E.cast(ctx, ref))
    wp = torch.randn(32, 5, 4, 4, device=dev)                          # DiT patch embedding: K order (py, px, c)
    assert torch.equal(E.pack_conv(ctx, wp), E.cast(ctx, wp.permute(0, 2, 3, 1).reshape(32, -1).contiguous()))
    inner, C = 256, 24
    gw, gb = torch.randn(2 * inner, C, device=dev), torch.randn(2 * inner, device=dev)
    rw = torch.cat([gw[:inner].reshape(inner // 128, 128, C), gw[inner:].reshape(inner // 128, 128, C)], 1).reshape(2 * inner, C)
    rb = torch.cat([gb[:inner].reshape(-1, 128), gb[inner:].reshape(-1, 128)], 1).reshape(-1)
    pw, pb = E.pack_geglu(ctx, gw, gb)
    assert torch.equal(pw, E.cast(ctx, rw.contiguous())) and torch.equal(pb, rb)
    wu = torch.randn(16, 12, 3, 3, device=dev)
    for got, want in zip(E.pack_upsample(ctx, wu), E.fold_upsample_weights(wu)):
        assert got.dtype == dt and torch.equal(got, E.cast(ctx, want))
    a, b = torch.randn(7, 10, device=dev), torch.randn(5, 10, device=dev)
    assert torch.equal(E.pack_rows(ctx, [a, b], 0), E.cast(ctx, torch.cat([a, b], 0)))
    c = torch.randn(7, 6, device=dev)
    assert torch.equal(E.pack_rows(ctx, [a, c], 1), E.cast(ctx, torch.cat([a, c], 1).contiguous()))
