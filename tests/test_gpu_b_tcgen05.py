"""GPU parity of the tcgen05/TMEM/TMA implicit-GEMM engine.

Checked three ways on the same bf16 inputs: against torch fp32 math (conv2d / matmul on the bf16-rounded operands),
against the independent CUDA-core SIMT engine of this library, and through fused-epilogue identities
(fused eps-MSE == unfused reduction of the written prediction).  Tolerance: bf16 output rounding (4e-3 relative
per element -> 1e-2 bound on the Frobenius-relative error; fp32 outputs must agree to 2e-5)."""
import pytest
import torch
import torch.nn.functional as F

from helpers import conv_ref, rel_err

pytestmark = pytest.mark.gpu


def _ctx(dev, engine=0):
    from dcb200 import engine as E
    return E.Ctx(device=dev, precision="bf16", engine=engine)


def _bf(*shape, dev, scale=1.0):
    return (torch.randn(*shape, device=dev) * scale).to(torch.bfloat16)


@pytest.mark.parametrize("M,K,N", [(128, 64, 128), (300, 128, 128), (64, 256, 48), (1000, 512, 256), (40, 128, 1536),
                                    (128 * 160, 128, 128), (333, 192, 3)])
def test_tc_linear(dev, M, K, N):
    from dcb200 import _lib as L
    from dcb200 import engine as E
    torch.manual_seed(0)
    x, w, b = _bf(M, K, dev=dev), _bf(N, K, dev=dev, scale=0.1), torch.randn(N, device=dev)
    ref = x.float() @ w.float().t() + b
    out32 = E.linear(_ctx(dev), x, w, N, bias=b, out_dtype=torch.float32)
    assert rel_err(out32, ref) < 2e-5, "tcgen05 fp32-out vs torch"
    simt = E.linear(_ctx(dev, L.ENGINE_SIMT), x, w, N, bias=b, out_dtype=torch.float32)
    assert rel_err(out32, simt) < 2e-5, "tcgen05 vs SIMT engine"
    out16 = E.linear(_ctx(dev), x, w, N, bias=b)
    assert out16.dtype == torch.bfloat16 and rel_err(out16, ref) < 1e-2


@pytest.mark.parametrize("NB,H,W,Ci,Co,stride", [
    (2, 32, 32, 64, 128, 1),     # bw=32 bh=4
    (3, 8, 8, 128, 256, 1),      # two samples per M tile (bn=2), ragged last tile
    (9, 4, 4, 64, 512, 1),       # bn=8, N split over tiles
    (1, 16, 256, 64, 64, 1),     # OW > 128: two x tiles per row
    (2, 128, 128, 64, 128, 1),   # full-res geometry of unet-128 (bw=128)
    (2, 32, 32, 64, 128, 2),     # Downsample2D: 5-D (2C, W/2, 2, H/2, N) view
    (5, 8, 8, 256, 256, 2),      # stride 2 with bn>1 at the output (4x4 -> bn=8)
    (2, 16, 16, 128, 3, 1),      # conv_out-like: N=3 (BN=16, masked columns)
])
def test_tc_conv3x3(dev, NB, H, W, Ci, Co, stride):
    from dcb200 import _lib as L
    from dcb200 import engine as E
    torch.manual_seed(0)
    x = _bf(NB, H, W, Ci, dev=dev)
    w = (torch.randn(Co, Ci, 3, 3, device=dev) * 0.05).to(torch.bfloat16)
    b = torch.randn(Co, device=dev)
    wp = w.permute(0, 2, 3, 1).reshape(Co, -1).contiguous()
    OH, OW = H // stride, W // stride
    ref = conv_ref(x.float(), w.float(), b, stride).reshape(-1, Co)
    out = E.gemm(_ctx(dev), E.conv3x3_segs(x, Ci, H, W, stride), wp, Co, NB, OH, OW, bias=b, out_dtype=torch.float32)
    assert rel_err(out, ref) < 2e-5
    simt = E.gemm(_ctx(dev, L.ENGINE_SIMT), E.conv3x3_segs(x, Ci, H, W, stride), wp, Co, NB, OH, OW, bias=b,
                  out_dtype=torch.float32)
    assert rel_err(out, simt) < 2e-5


def test_tc_resnet_tail_concat_shortcut(dev):
    """conv2 + 1x1 conv_shortcut over cat([h, skip]) as extra K segments of ONE GEMM (ResnetBlock2D in an up block)."""
    from dcb200 import engine as E
    torch.manual_seed(0)
    NB, H, W, C0, C1, Co = 2, 16, 16, 256, 128, 128
    a2, x0, x1 = _bf(NB, H, W, Co, dev=dev), _bf(NB, H, W, C0, dev=dev), _bf(NB, H, W, C1, dev=dev)
    w2 = (torch.randn(Co, Co, 3, 3, device=dev) * 0.05).to(torch.bfloat16)
    ws = (torch.randn(Co, C0 + C1, 1, 1, device=dev) * 0.05).to(torch.bfloat16)
    b = torch.randn(Co, device=dev)
    wp = torch.cat([w2.permute(0, 2, 3, 1).reshape(Co, -1), ws.reshape(Co, -1)], 1).contiguous()
    segs = E.conv3x3_segs(a2, Co, H, W) + [E.seg(x0, C0, H, W), E.seg(x1, C1, H, W)]
    out = E.gemm(_ctx(dev), segs, wp, Co, NB, H, W, bias=b, out_dtype=torch.float32)
    ref = conv_ref(a2.float(), w2.float(), b) + conv_ref(torch.cat([x0, x1], -1).float(), ws.float(), None, 1, 0)
    assert rel_err(out.reshape(NB, H, W, Co), ref) < 2e-5


def test_tc_epilogue_flags(dev):
    from dcb200 import _lib as L
    from dcb200 import engine as E
    torch.manual_seed(0)
    M, K, N, rpg = 512, 128, 256, 128
    x, w, b = _bf(M, K, dev=dev), _bf(N, K, dev=dev, scale=0.1), torch.randn(N, device=dev)
    rv, gate = torch.randn(5, N, device=dev), torch.randn(4, N, device=dev)
    res = _bf(M, N, dev=dev)
    idx = torch.tensor([4, 0, 2, 1], device=dev, dtype=torch.int32)
    out = E.linear(_ctx(dev), x, w, N, bias=b, rowvec=rv, rowvec_ld=N, rowvec_idx=idx, rows_per_group=rpg, gate=gate,
                   gate_ld=N, residual=res, res_ld=N, act=L.ACT_GELU_TANH, act_post=L.ACT_SILU, out_dtype=torch.float32)
    grp = torch.arange(M, device=dev) // rpg
    v = x.float() @ w.float().t() + b + rv[idx.long()][grp]
    ref = F.silu(F.gelu(v, approximate="tanh") * gate[grp] + res.float())
    assert rel_err(out, ref) < 2e-5
    # gathered residual rows (DiT label embedding) and modulo residual (DiT positional embedding)
    table = _bf(7, N, dev=dev)
    lab = torch.randint(0, 7, (M,), device=dev, dtype=torch.int32)
    out = E.linear(_ctx(dev), x, w, N, residual=table, res_ld=N, res_idx=lab, out_dtype=torch.float32)
    assert rel_err(out, x.float() @ w.float().t() + table.float()[lab.long()]) < 2e-5
    pos = _bf(64, N, dev=dev)
    out = E.linear(_ctx(dev), x, w, N, residual=pos, res_ld=N, res_mod=64, out_dtype=torch.float32)
    assert rel_err(out, x.float() @ w.float().t() + pos.float().repeat(M // 64, 1)) < 2e-5


def test_tc_geglu(dev):
    from dcb200 import _lib as L
    from dcb200 import engine as E
    torch.manual_seed(0)
    M, K, inner = 200, 256, 1024
    x = _bf(M, K, dev=dev)
    w, b = _bf(2 * inner, K, dev=dev, scale=0.1), torch.randn(2 * inner, device=dev)
    wp = torch.cat([w[:inner].reshape(-1, 128, K), w[inner:].reshape(-1, 128, K)], 1).reshape(2 * inner, K).contiguous()
    bp = torch.cat([b[:inner].reshape(-1, 128), b[inner:].reshape(-1, 128)], 1).reshape(-1).contiguous()
    out = E.linear(_ctx(dev), x, wp, 2 * inner, bias=bp, act=L.ACT_GEGLU, out_dtype=torch.float32)
    h, g = (x.float() @ w.float().t() + b).chunk(2, -1)
    assert out.shape == (M, inner) and rel_err(out, h * F.gelu(g)) < 2e-5


@pytest.mark.parametrize("v_param", [False, True])
def test_tc_fused_mse_equals_unfused(dev, v_param):
    """conv_out with the eps-MSE epilogue (pred never written) == ||scale*pred - target||^2 of the written pred
    (diffusion_classifier.py:706-711), per sample, with target shared by `div` class slots."""
    from dcb200 import engine as E
    torch.manual_seed(0)
    S, div, H, W, Ci, Co = 6, 2, 32, 32, 128, 3
    a = _bf(S, H, W, Ci, dev=dev)
    w = (torch.randn(Co, Ci, 3, 3, device=dev) * 0.05).to(torch.bfloat16)
    b = torch.randn(Co, device=dev)
    wp = w.permute(0, 2, 3, 1).reshape(Co, -1).contiguous()
    tgt = torch.randn(S // div, H * W, Co, device=dev)
    scale = torch.rand(S, device=dev) if v_param else None
    err = torch.empty(S, device=dev)
    ctx = _ctx(dev)
    E.gemm(ctx, E.conv3x3_segs(a, Ci, H, W), wp, Co, S, H, W, bias=b, want_out=False,
           mse=dict(target=tgt, scale=scale, div=div, ld=Co, err=err))
    pred = conv_ref(a.float(), w.float(), b).reshape(S, H * W, Co)
    sc = scale.view(-1, 1, 1) if v_param else 1.0
    ref = ((sc * pred - tgt.repeat_interleave(div, 0)) ** 2).sum((1, 2))
    assert rel_err(err, ref) < 2e-5
    err2 = torch.empty(S, device=dev)
    E.gemm(ctx, E.conv3x3_segs(a, Ci, H, W), wp, Co, S, H, W, bias=b, want_out=False,
           mse=dict(target=tgt, scale=scale, div=div, ld=Co, err=err2))
    assert torch.equal(err, err2), "fixed-order reduction must be run-to-run deterministic"


def test_tc_persistent_many_tiles_deep_k(dev):
    """more tiles than SMs (persistent loop, TMEM double buffering) and K deep enough to wrap the smem ring often."""
    from dcb200 import engine as E
    torch.manual_seed(0)
    NB, H, W, Ci, Co = 4, 64, 64, 256, 256   # 128 M tiles... x1 N tile; K = 2304 (36 K blocks)
    x = _bf(NB, H, W, Ci, dev=dev)
    w = (torch.randn(Co, Ci, 3, 3, device=dev) * 0.03).to(torch.bfloat16)
    wp = w.permute(0, 2, 3, 1).reshape(Co, -1).contiguous()
    out = E.gemm(_ctx(dev), E.conv3x3_segs(x, Ci, H, W), wp, Co, NB, H, W, out_dtype=torch.float32)
    assert rel_err(out, conv_ref(x.float(), w.float(), None).reshape(-1, Co)) < 2e-5
    M, K, N = 128 * 400, 64, 128             # 400 tiles on 148 SMs, single K block each
    a, wl = _bf(M, K, dev=dev), _bf(N, K, dev=dev, scale=0.1)
    out = E.linear(_ctx(dev), a, wl, N, out_dtype=torch.float32)
    assert rel_err(out, a.float() @ wl.float().t()) < 2e-5


# ---- 256-pixel CTA kernel (gemm_tc2.cu): split A/B rings, two accumulators, x-halo reuse -------------------------
@pytest.mark.parametrize("NB,H,W,Ci,Co,stride,res", [
    (5, 128, 128, 64, 128, 1, False),    # x-halo mode, OW=128 (sub-tiles = two image rows)
    (5, 64, 256, 128, 128, 1, True),     # x-halo mode, OW=256 (sub-tiles = two halves of a row) + residual prefetch
    (5, 128, 128, 128, 64, 1, False),    # x-halo, BN=64
    (20, 64, 64, 64, 128, 1, True),      # tap mode (bw=64, bh=2)
    (10, 128, 128, 64, 128, 2, False),   # tap mode, stride 2
    (37, 32, 32, 64, 128, 1, False),     # tap mode, odd number of 128-row sub-tiles (tail box fully out of range)
])
def test_tc2_conv3x3(dev, NB, H, W, Ci, Co, stride, res):
    from dcb200 import _lib as L
    from dcb200 import engine as E
    torch.manual_seed(0)
    x = _bf(NB, H, W, Ci, dev=dev)
    w = (torch.randn(Co, Ci, 3, 3, device=dev) * 0.05).to(torch.bfloat16)
    b = torch.randn(Co, device=dev)
    wp = w.permute(0, 2, 3, 1).reshape(Co, -1).contiguous()
    OH, OW = H // stride, W // stride
    r = _bf(NB * OH * OW, Co, dev=dev) if res else None
    ref = conv_ref(x.float(), w.float(), b, stride).reshape(-1, Co)
    if res:
        ref = ref + r.float()
    kw = dict(bias=b, residual=r, res_ld=Co) if res else dict(bias=b)
    out = E.gemm(_ctx(dev), E.conv3x3_segs(x, Ci, H, W, stride), wp, Co, NB, OH, OW, **kw)
    assert out.dtype == torch.bfloat16 and rel_err(out, ref) < 5e-3
    simt = E.gemm(_ctx(dev, L.ENGINE_SIMT), E.conv3x3_segs(x, Ci, H, W, stride), wp, Co, NB, OH, OW, **kw)
    assert rel_err(out, simt.float()) < 1e-3      # same bf16 inputs, both round the fp32 result to bf16 once


def test_tc2_halo_with_shortcut_segments_and_rowvec(dev):
    """full-resolution up-block resnet tail: 3x3 conv (x-halo) + 1x1 shortcut over cat([h, skip]) as plain tap segments,
    and conv1-style per-sample row vector, in the 256-pixel kernel."""
    from dcb200 import engine as E
    torch.manual_seed(0)
    NB, H, W, C0, C1, Co = 5, 128, 128, 128, 64, 128
    a2, x0, x1 = _bf(NB, H, W, Co, dev=dev), _bf(NB, H, W, C0, dev=dev), _bf(NB, H, W, C1, dev=dev)
    w2 = (torch.randn(Co, Co, 3, 3, device=dev) * 0.05).to(torch.bfloat16)
    ws = (torch.randn(Co, C0 + C1, 1, 1, device=dev) * 0.05).to(torch.bfloat16)
    b, rv = torch.randn(Co, device=dev), torch.randn(NB, Co, device=dev)
    wp = torch.cat([w2.permute(0, 2, 3, 1).reshape(Co, -1), ws.reshape(Co, -1)], 1).contiguous()
    segs = E.conv3x3_segs(a2, Co, H, W) + [E.seg(x0, C0, H, W), E.seg(x1, C1, H, W)]
    out = E.gemm(_ctx(dev), segs, wp, Co, NB, H, W, bias=b, rowvec=rv, rowvec_ld=Co, rows_per_group=H * W)
    ref = conv_ref(a2.float(), w2.float(), b) + conv_ref(torch.cat([x0, x1], -1).float(), ws.float(), None, 1, 0)
    ref = ref + rv[:, None, None, :]
    assert rel_err(out.reshape(NB, H, W, Co), ref) < 5e-3


def test_tc2_linear_odd_tiles(dev):
    from dcb200 import engine as E
    torch.manual_seed(0)
    M, K, N = 128 * 601 + 77, 128, 128
    x, w, b = _bf(M, K, dev=dev), _bf(N, K, dev=dev, scale=0.1), torch.randn(N, device=dev)
    r = _bf(M, N, dev=dev)
    out = E.linear(_ctx(dev), x, w, N, bias=b, residual=r, res_ld=N)
    assert rel_err(out, x.float() @ w.float().t() + b + r.float()) < 5e-3


@pytest.mark.parametrize("U,rep,H,W,Ci,Co,big", [
    (3, 2, 16, 16, 128, 128, False),    # gemm_tc (few tiles), 1x1 + 3x3 over a per-unit source
    (5, 3, 128, 128, 64, 128, True),    # gemm_tc2 incl. the x-halo path: many tiles
    (2, 10, 32, 32, 64, 64, False),     # CIFAR-like: 10 classes per unit
])
def test_tc_unit_shared_sources(dev, U, rep, H, W, Ci, Co, big):
    """dcb_seg.nb_div: a per-(image, timestep) unit tensor read by all of the unit's class-conditional samples equals
    the GEMM over the materialised repeat_interleave'd tensor, bit for bit, in both tcgen05 kernels and the SIMT engine."""
    from dcb200 import _lib as L
    from dcb200 import engine as E
    torch.manual_seed(0)
    S = U * rep
    xu = _bf(U, H, W, Ci, dev=dev)
    xs = xu.repeat_interleave(rep, 0).contiguous()
    ys = _bf(S, H, W, Ci, dev=dev)                                # a genuinely per-sample second source
    w = (torch.randn(Co, 10 * Ci, device=dev) * 0.05).to(torch.bfloat16)
    b = torch.randn(Co, device=dev)
    for eng in (0, L.ENGINE_SIMT):
        segs_u = E.conv3x3_segs(ys, Ci, H, W) + [E.seg(xu, Ci, H, W, nb_div=rep)]
        segs_s = E.conv3x3_segs(ys, Ci, H, W) + [E.seg(xs, Ci, H, W)]
        o_u = E.gemm(_ctx(dev, eng), segs_u, w, Co, S, H, W, bias=b)
        o_s = E.gemm(_ctx(dev, eng), segs_s, w, Co, S, H, W, bias=b)
        assert torch.equal(o_u, o_s), f"1x1 unit source, engine {eng}"
        w9 = w[:, :9 * Ci].contiguous()
        o_u = E.gemm(_ctx(dev, eng), E.conv3x3_segs(xu, Ci, H, W, nb_div=rep), w9, Co, S, H, W, bias=b)
        o_s = E.gemm(_ctx(dev, eng), E.conv3x3_segs(xs, Ci, H, W), w9, Co, S, H, W, bias=b)
        assert torch.equal(o_u, o_s), f"3x3 unit source, engine {eng}"
    # residual rows gathered from the unit tensor (attn1 out-projection of a shared attention core)
    HW = H * W
    m = torch.arange(S * HW, device=dev)
    ridx = ((m // (HW * rep)) * HW + m % HW).to(torch.int32)
    r_u = _bf(U * HW, Co, dev=dev)
    w1 = w[:, :Ci].contiguous()
    o_u = E.gemm(_ctx(dev), [E.seg(xu, Ci, H, W, nb_div=rep)], w1, Co, S, H, W, bias=b, residual=r_u, res_ld=Co, res_idx=ridx)
    o_s = E.gemm(_ctx(dev), [E.seg(xs, Ci, H, W)], w1, Co, S, H, W, bias=b,
                 residual=r_u.reshape(U, HW, Co).repeat_interleave(rep, 0).reshape(-1, Co).contiguous(), res_ld=Co)
    assert torch.equal(o_u, o_s)


def test_tc_unit_source_needs_single_sample_tiles(dev):
    from dcb200 import _lib as L
    from dcb200 import engine as E
    x = _bf(2, 8, 8, 64, dev=dev)   # 64 pixels per sample: a 128-row tile spans two samples
    w = _bf(64, 64, dev=dev)
    with pytest.raises(L.DcbError):
        E.gemm(_ctx(dev, L.ENGINE_TCGEN05), [E.seg(x, 64, 8, 8, nb_div=2)], w, 64, 4, 8, 8)


@pytest.mark.parametrize("NB,H,W,Ci,Co,res", [(3, 16, 16, 128, 128, True), (2, 128, 128, 64, 128, False),
                                              (5, 32, 32, 64, 256, True), (9, 128, 128, 64, 64, False)])
def test_tc_epilogue_groupnorm_tile_statistics(dev, NB, H, W, Ci, Co, res):
    """dcb_gemm_desc.gn_part: the staged epilogue's per-tile (sum, sumsq) of the bf16 values it stores, and GroupNorm
    fed from them (dcb_groupnorm_stats_from_tiles) vs GroupNorm with its own statistics pass over the same tensor."""
    from dcb200 import engine as E
    torch.manual_seed(0)
    ctx = _ctx(dev)
    HW = H * W
    x = _bf(NB, H, W, Ci, dev=dev)
    w = (torch.randn(Co, 9 * Ci, device=dev) * 0.05).to(torch.bfloat16)
    b = torch.randn(Co, device=dev)
    r = _bf(NB * HW, Co, dev=dev) if res else None
    out, part = E.gemm(ctx, E.conv3x3_segs(x, Ci, H, W), w, Co, NB, H, W, bias=b, residual=r, res_ld=Co, gn_stats=True)
    plain = E.gemm(ctx, E.conv3x3_segs(x, Ci, H, W), w, Co, NB, H, W, bias=b, residual=r, res_ld=Co)
    assert torch.equal(out, plain), "asking for statistics must not change the output"
    assert part is not None and part.shape == (NB * HW // 128, Co, 2)
    tps = HW // 128
    # per-sample, per-channel totals (tile -> pixel assignment is a box, so compare after summing a sample's tiles)
    got = part.reshape(NB, tps, Co, 2).double().sum(1)
    o = out.reshape(NB, HW, Co).double()
    assert rel_err(got[..., 0], o.sum(1)) < 1e-5 and rel_err(got[..., 1], (o * o).sum(1)) < 1e-5
    g, be = torch.randn(Co, device=dev), torch.randn(Co, device=dev)
    a = E.groupnorm(ctx, out, Co, None, 0, NB, HW, g, be, 1e-5, True, st0=part)
    ref = E.groupnorm(ctx, out, Co, None, 0, NB, HW, g, be, 1e-5, True)
    assert rel_err(a, ref) < 2e-3 and (a.float() - ref.float()).abs().max() < 0.07   # bf16 ulps where rounding flips
    # concatenated sources, one of them per unit
    rep = 1 if NB % 3 else 3
    sk = _bf(NB // rep, H, W, Ci, dev=dev)
    sk_out, sk_part = E.gemm(ctx, [E.seg(sk, Ci, H, W)], w[:, :Ci].contiguous(), Co, NB // rep, H, W, bias=b, gn_stats=True)
    g2, b2 = torch.randn(2 * Co, device=dev), torch.randn(2 * Co, device=dev)
    a = E.groupnorm(ctx, out, Co, sk_out, Co, NB, HW, g2, b2, 1e-5, True, div1=rep, st0=part, st1=sk_part)
    ref = E.groupnorm(ctx, out, Co, sk_out, Co, NB, HW, g2, b2, 1e-5, True, div1=rep)
    assert rel_err(a, ref) < 2e-3


@pytest.mark.parametrize("NB,H,W,C", [(2, 64, 64, 128), (3, 16, 16, 256), (5, 8, 8, 128), (3, 4, 4, 64), (130, 32, 32, 64)])
def test_upsample_conv_folded_phases(dev, NB, H, W, C):
    """Upsample2D (a5.4: nearest 2x then conv3x3 pad 1) as four 2x2-tap phase GEMMs over the low-resolution input
    (dcb_gemm_desc.up_phase) vs torch interpolate + conv2d, vs the unfolded upsample2x + 3x3 GEMM, on the tcgen05 and
    SIMT engines; tile statistics of the interleaved output vs the written tensor."""
    from dcb200 import _lib as L
    from dcb200 import engine as E
    torch.manual_seed(0)
    x = _bf(NB, H, W, C, dev=dev)
    w32 = torch.randn(C, C, 3, 3, device=dev) * 0.05
    b = torch.randn(C, device=dev)
    wph32 = E.fold_upsample_weights(w32)
    wph = [p.to(torch.bfloat16) for p in wph32]
    up = F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest")
    # exact identity in real arithmetic: checked in fp32 with the un-rounded folded weights on the CUDA-core engine
    ref32 = F.conv2d(up, w32, b, padding=1).permute(0, 2, 3, 1).reshape(-1, C)
    c32 = E.Ctx(device=dev, precision="fp32")
    o32, _ = E.upsample_conv(c32, x.float().reshape(-1, C).contiguous(), wph32, b, C, C, NB, H, W)
    assert rel_err(o32, ref32) < 1e-5
    # bf16: reference = the same bf16-rounded phase weights applied by torch (isolates the kernel from weight rounding)
    ref = torch.empty(NB, 2 * H, 2 * W, C, device=dev)
    xp = F.pad(x.float().permute(0, 3, 1, 2), (1, 1, 1, 1))
    for a in range(2):
        for bb in range(2):
            wk = wph[2 * a + bb].float().reshape(C, 2, 2, C).permute(0, 3, 1, 2)
            y = F.conv2d(xp[:, :, a:a + H + 1, bb:bb + W + 1], wk, b)
            ref[:, a::2, bb::2, :] = y.permute(0, 2, 3, 1)
    ref = ref.reshape(-1, C)
    assert rel_err(ref, ref32) < 1e-2          # folded-then-rounded weights vs the 3x3 conv: bf16 weight rounding only
    ctx = _ctx(dev)
    out, st = E.upsample_conv(ctx, x.reshape(-1, C), wph, b, C, C, NB, H, W)
    assert out.dtype == torch.bfloat16 and rel_err(out, ref) < 4e-3
    simt, _ = E.upsample_conv(_ctx(dev, L.ENGINE_SIMT), x.reshape(-1, C), wph, b, C, C, NB, H, W)
    assert rel_err(out, simt) < 4e-3
    if H * W % 128 == 0:
        assert st is not None and st.shape == (NB * 4 * H * W // 128, C, 2)
        got = st.reshape(NB, -1, C, 2).double().sum(1)
        o = out.reshape(NB, 4 * H * W, C).double()
        assert rel_err(got[..., 0], o.sum(1)) < 1e-5 and rel_err(got[..., 1], (o * o).sum(1)) < 1e-5
    else:
        assert st is None


@pytest.mark.parametrize("S,div,H,W,Ci,Co", [(5, 1, 128, 128, 128, 3), (6, 2, 128, 128, 64, 12), (3, 1, 64, 256, 128, 40)])
def test_tc2_halo_fused_mse_is_bit_identical_to_tap_kernel(dev, S, div, H, W, Ci, Co):
    """conv_out at full resolution runs in gemm_tc2 (x-halo boxes, direct eps-MSE epilogue): same K-block order and the
    same summation order as gemm_tc_kernel with nine separately loaded taps => bit-identical per-sample errors; both
    equal the torch fp32 reduction of the written prediction."""
    from dcb200 import _lib as L
    from dcb200 import engine as E
    torch.manual_seed(0)
    a = _bf(S, H, W, Ci, dev=dev)
    w = (torch.randn(Co, Ci, 3, 3, device=dev) * 0.05).to(torch.bfloat16)
    b = torch.randn(Co, device=dev)
    wp = w.permute(0, 2, 3, 1).reshape(Co, -1).contiguous()
    tgt = torch.randn(S // div, H * W, Co, device=dev)
    scale = torch.rand(S, device=dev)
    ctx = _ctx(dev)

    def run():
        err = torch.empty(S, device=dev)
        E.gemm(ctx, E.conv3x3_segs(a, Ci, H, W), wp, Co, S, H, W, bias=b, want_out=False,
               mse=dict(target=tgt, scale=scale, div=div, ld=Co, err=err))
        return err

    e_halo = run()
    with L.knob("NO_TC2_MSE"):
        e_tap = run()
    assert torch.equal(e_halo, e_tap)
    pred = conv_ref(a.float(), w.float(), b).reshape(S, H * W, Co)
    ref = ((scale.view(-1, 1, 1) * pred - tgt.repeat_interleave(div, 0)) ** 2).sum((1, 2))
    assert rel_err(e_halo, ref) < 2e-5


@pytest.mark.parametrize("NB,H,W,Ci,Co,extra", [
    (20, 64, 64, 128, 128, "res"),        # OW = 64: box [64 px x 6 rows], residual prefetch
    (40, 32, 32, 64, 256, "rowvec"),      # OW = 32: box [32 px x 10 rows], two N tiles, conv1-style row vector
    (160, 16, 16, 128, 128, "shortcut"),  # OW = 16: box [16 px x 18 rows] + 1x1 shortcut segments as plain taps
    (12, 64, 64, 64, 64, "unit"),         # per-unit source (dcb_seg.nb_div = 2) and BN = 64
])
def test_tc2_yhalo_conv3x3(dev, NB, H, W, Ci, Co, extra):
    """y-halo mode of gemm_tc2 (rows narrower than a tile: one [OW px x (2 bh + 2) rows] box per (kx, channel block) serves
    the three ky taps of both sub-tiles) vs torch, vs the SIMT engine and -- bit for bit, same K-block order -- vs the same
    kernel with nine separately loaded taps (DCB_TC2_NO_YHALO=1) and vs gemm_tc_kernel (DCB_NO_TC2=1)."""
    import os
    from dcb200 import _lib as L
    from dcb200 import engine as E
    torch.manual_seed(0)
    div = 2 if extra == "unit" else 1
    x = _bf(NB // div, H, W, Ci, dev=dev)
    w = (torch.randn(Co, Ci, 3, 3, device=dev) * 0.05).to(torch.bfloat16)
    b = torch.randn(Co, device=dev)
    wp = w.permute(0, 2, 3, 1).reshape(Co, -1).contiguous()
    segs = E.conv3x3_segs(x, Ci, H, W, nb_div=div)
    xr = x.float().repeat_interleave(div, 0)
    ref = conv_ref(xr, w.float(), b).reshape(-1, Co)
    kw = dict(bias=b)
    if extra == "res":
        r = _bf(NB * H * W, Co, dev=dev)
        kw.update(residual=r, res_ld=Co)
        ref = ref + r.float()
    elif extra == "rowvec":
        rv = torch.randn(NB, Co, device=dev)
        kw.update(rowvec=rv, rowvec_ld=Co, rows_per_group=H * W)
        ref = (ref.reshape(NB, H * W, Co) + rv[:, None, :]).reshape(-1, Co)
    elif extra == "shortcut":
        x0 = _bf(NB, H, W, 64, dev=dev)
        ws = (torch.randn(Co, 64, 1, 1, device=dev) * 0.05).to(torch.bfloat16)
        wp = torch.cat([wp, ws.reshape(Co, -1)], 1).contiguous()
        segs = segs + [E.seg(x0, 64, H, W)]
        ref = ref + conv_ref(x0.float(), ws.float(), None, 1, 0).reshape(-1, Co)
    out, st = E.gemm(_ctx(dev), segs, wp, Co, NB, H, W, gn_stats=True, **kw)
    assert out.dtype == torch.bfloat16 and rel_err(out, ref) < 5e-3
    simt = E.gemm(_ctx(dev, L.ENGINE_SIMT), segs, wp, Co, NB, H, W, **kw)
    assert rel_err(out, simt.float()) < 1e-3
    for knob in ("TC2_NO_YHALO", "NO_TC2"):
        with L.knob(knob):
            other, st2 = E.gemm(_ctx(dev), segs, wp, Co, NB, H, W, gn_stats=True, **kw)
        assert torch.equal(out, other), knob
        assert torch.equal(st, st2), knob


@pytest.mark.parametrize("M,K,N,extra", [
    (128 * 200, 768, 768, "gate_res"),        # DiT attention out-projection: adaLN gate + residual
    (128 * 203 + 50, 768, 2304, "bias"),      # QKV; odd number of 128-row sub-tiles + ragged last tile
    (128 * 240, 768, 3072, "gelu"),           # DiT FFN1: tanh-GELU in the direct epilogue (one-MUFU form, as staged)
    (128 * 400, 3072, 768, "gate_res"),       # DiT FFN2 (deep K)
    (128 * 300, 512, 1536, "rowvec"),         # U-Net QKV shape with a per-group row vector
])
def test_tc2_wide_256x256_tiles(dev, M, K, N, extra):
    """experimental wide mode of gemm_tc2 (DCB_TC2_WIDE=1; BN = 256: two 128 x 256 accumulators, single TMEM stage, direct
    epilogue -- measured slower than the default, see gemm_tc.cu) vs torch fp32 math and, bit for bit, vs the default
    128-wide staged path: the direct and staged epilogues perform the same arithmetic in the same order."""
    import os
    from dcb200 import _lib as L
    from dcb200 import engine as E
    torch.manual_seed(0)
    x, w, b = _bf(M, K, dev=dev), _bf(N, K, dev=dev, scale=0.05), torch.randn(N, device=dev)
    rpg = 128 * 25
    ngrp = (M + rpg - 1) // rpg
    kw = dict(bias=b)
    ref = x.float() @ w.float().t() + b
    grp = torch.arange(M, device=dev) // rpg
    if extra == "gate_res":
        gate, res = torch.randn(ngrp, N, device=dev), _bf(M, N, dev=dev)
        kw.update(gate=gate, gate_ld=N, rows_per_group=rpg, residual=res, res_ld=N)
        ref = ref * gate[grp] + res.float()
    elif extra == "gelu":
        kw.update(act=L.ACT_GELU_TANH)
        ref = F.gelu(ref, approximate="tanh")
    elif extra == "rowvec":
        rv = torch.randn(ngrp, N, device=dev)
        kw.update(rowvec=rv, rowvec_ld=N, rows_per_group=rpg)
        ref = ref + rv[grp]
    narrow = E.linear(_ctx(dev), x, w, N, **kw)
    assert narrow.dtype == torch.bfloat16 and rel_err(narrow, ref) < 6e-3
    n0 = L.launch_count()
    with L.knob("TC2_WIDE"):
        out = E.linear(_ctx(dev), x, w, N, **kw)
    assert L.launch_count() == n0 + 1 and torch.equal(out, narrow)


def _tile_stats(x):
    """per-128-row-tile, per-column (sum, sum of squares) of a [rows, C] bf16 tensor, as a producer GEMM's epilogue writes"""
    f = x.float().reshape(-1, 128, x.shape[-1])
    return torch.stack([f.sum(1), (f * f).sum(1)], -1).contiguous()


@pytest.mark.parametrize("NB,H,W,C0,C1,div1,N,extra", [
    (5, 128, 128, 128, 0, 1, 128, "rowvec"),       # resnet conv1, down path: single source, time-embedding row vector
    (6, 128, 128, 128, 128, 2, 128, "rowvec"),     # resnet conv1, up path: cat([h, per-unit skip]) -> K = 9 * 256 (dominant conv)
    (5, 128, 128, 128, 0, 1, 128, "residual"),     # resnet conv2 with identity shortcut
    (5, 128, 128, 128, 0, 1, 128, "shortcut"),     # resnet conv2 + 1x1 conv_shortcut over the raw (concatenated) input
    (6, 128, 128, 128, 0, 1, 3, "mse"),            # conv_norm_out + conv_out + fused eps-MSE (direct epilogue, 4 slots)
    (10, 32, 256, 64, 64, 1, 64, "nosilu"),        # 256-pixel rows (two tiles per row), BN = 64, affine only
])
def test_gn_fused_into_conv_operand_is_bit_identical(dev, NB, H, W, C0, C1, div1, N, extra):
    """XF variant of gemm_tc2 (transform warps rewrite every x-halo box in place: y = a[n,c] x + b[n,c], SiLU) vs the
    separate gn_apply pass + plain conv: same coefficients, same fp32 arithmetic, same bf16 rounding, same K order =>
    outputs and next-layer tile statistics are bit-identical; and within bf16 tolerance of torch's group_norm + conv2d."""
    from dcb200 import _lib as L
    from dcb200 import engine as E
    torch.manual_seed(0)
    ctx = _ctx(dev)
    x0 = _bf(NB * H * W, C0, dev=dev) * 1.5 + 0.3
    x1 = (_bf((NB // div1) * H * W, C1, dev=dev) * 0.7 - 0.2) if C1 else None
    Ct = C0 + C1
    gamma, beta = torch.randn(Ct, device=dev), torch.randn(Ct, device=dev) * 0.3
    w = _bf(N, Ct, 3, 3, dev=dev, scale=0.05)
    wp = w.permute(0, 2, 3, 1).reshape(N, -1).contiguous()
    b = torch.randn(N, device=dev)
    st0, st1 = _tile_stats(x0), (_tile_stats(x1) if C1 else None)
    silu = extra != "nosilu"
    kw = dict(bias=b)
    extra_segs = ()
    ws = None
    if extra == "rowvec":
        rv = torch.randn(NB, N, device=dev)
        kw.update(rowvec=rv, rowvec_ld=N, rows_per_group=H * W, gn_stats=True)
    elif extra == "residual":
        res = _bf(NB * H * W, N, dev=dev)
        kw.update(residual=res, res_ld=N, gn_stats=True)
    elif extra == "shortcut":
        xs = _bf(NB * H * W, 64, dev=dev)
        ws = _bf(N, 64, dev=dev, scale=0.05)
        wp = torch.cat([wp, ws], 1).contiguous()
        extra_segs = [E.seg(xs, 64, H, W)]
        kw.update(gn_stats=True)
    elif extra == "mse":
        tgt = torch.randn(NB * H * W, N, device=dev)
        kw.update(want_out=False)
    else:
        kw.update(gn_stats=True)

    def run(fuse):
        E.FUSE_GN = fuse
        if extra == "mse":
            err = torch.empty(NB, device=dev)
            kw["mse"] = dict(target=tgt, div=1, ld=N, err=err)
        n0 = L.launch_count()
        r = E.gn_conv3x3(ctx, x0, C0, x1, C1, NB, H, W, gamma, beta, 1e-5, silu, wp, N, div1=div1, st0=st0, st1=st1,
                         extra_segs=extra_segs, **kw)
        return (err,) if extra == "mse" else r, L.launch_count() - n0

    min_c, E.FUSE_GN_MSE = E.FUSE_GN_MIN_C, True
    E.FUSE_GN_MIN_C = 0           # whatever the product's policy knobs say, the kernel is tested on every layer shape
    try:
        fused, n_f = run(True)
        plain, n_p = run(False)
    finally:
        E.FUSE_GN, E.FUSE_GN_MIN_C, E.FUSE_GN_MSE = True, min_c, False
    # fused: coefficient kernel + conv (+ mse finalize); plain: statistics finalize + gn_apply + conv (+ mse finalize)
    assert n_f == n_p - 1, (n_f, n_p)
    for a, c in zip(fused, plain):
        assert torch.equal(a, c)
    # torch fp32 reference of the whole thing
    xin = x0.float().reshape(NB, H * W, C0)
    if C1:
        xin = torch.cat([xin, x1.float().reshape(NB // div1, H * W, C1).repeat_interleave(div1, 0)], -1)
    y = F.group_norm(xin.permute(0, 2, 1).reshape(NB, Ct, H, W), 32, gamma, beta, 1e-5)
    y = (F.silu(y) if silu else y).to(torch.bfloat16).float()
    ref = F.conv2d(y, w.float(), b, padding=1).permute(0, 2, 3, 1).reshape(NB * H * W, N)
    if extra == "rowvec":
        ref = ref + rv.repeat_interleave(H * W, 0)
    elif extra == "residual":
        ref = ref + res.float()
    elif extra == "shortcut":
        ref = ref + xs.float() @ ws.float().t()
    if extra == "mse":
        want = ((ref - tgt) ** 2).reshape(NB, -1).sum(1)
        assert rel_err(fused[0], want) < 2e-3
    else:
        assert rel_err(fused[0], ref) < 6e-3


@pytest.mark.parametrize("M,K,N,extra", [
    (40000, 768, 768, "plain"),          # M not a multiple of 256 (last pair: peer CTA fully out of range rows masked)
    (33000, 768, 2304, "bias"),          # DiT QKV
    (20480, 768, 3072, "gelu"),          # DiT FF1 (GELU-tanh epilogue)
    (20480, 3072, 768, "gate_res"),      # DiT FF2: adaLN gate + residual, K = 3072 (48 K blocks)
    (51200, 512, 512, "rowvec_idx_c_off"),   # U-Net attention out-projection: gathered row vector, K range inside a wider tensor
    (51200, 512, 4096, "geglu"),         # U-Net feed-forward: GEGLU, [128 value | 128 gate] weight rows per tile -> 2048 outputs
    (20480, 768, 768, "gate_res"),       # DiT attention out-projection (four epilogue groups: K <= 1024 with gate + residual)
])
def test_tc3_cta_pair_linear(dev, M, K, N, extra):
    """gemm_tc3_kernel (tcgen05 cta_group::2: a pair of CTAs computes a 256 x 256 tile, each loading its own A rows and half
    of W) vs the one-CTA kernel it replaces for these layers (DCB_KNOB_NO_TC3: bit for bit -- same K order, same staged
    epilogue arithmetic) and vs torch fp32 math."""
    from dcb200 import _lib as L
    from dcb200 import engine as E
    torch.manual_seed(0)
    ctx = _ctx(dev)
    Cw = K + 128 if extra == "rowvec_idx_c_off" else K
    x = _bf(M, Cw, dev=dev)
    w = _bf(N, K, dev=dev, scale=0.05)
    kw = {}
    xs = x[:, 64:64 + K] if extra == "rowvec_idx_c_off" else x
    ref = xs.float() @ w.float().t()
    rpg = 128 * 25
    ngrp = (M + rpg - 1) // rpg
    if extra != "plain":
        b = torch.randn(N, device=dev)
        kw["bias"] = b
        ref = ref + b
    if extra == "gelu":
        kw["act"] = L.ACT_GELU_TANH
        ref = F.gelu(ref, approximate="tanh")
    if extra == "geglu":          # packed rows: per 128 outputs [128 value rows | 128 gate rows]
        kw["act"] = L.ACT_GEGLU
        r4 = ref.reshape(M, N // 256, 2, 128)
        ref = (r4[:, :, 0] * F.gelu(r4[:, :, 1])).reshape(M, N // 2)
    if extra == "gate_res":
        gate = torch.randn(ngrp, N, device=dev)
        res = _bf(M, N, dev=dev)
        kw.update(gate=gate, gate_ld=N, rows_per_group=rpg, residual=res, res_ld=N)
        ref = ref * gate.repeat_interleave(rpg, 0)[:M] + res.float()
    if extra == "rowvec_idx_c_off":
        rv = torch.randn(7, N, device=dev)
        idx = torch.randint(0, 7, (ngrp,), device=dev, dtype=torch.int32)
        kw.update(rowvec=rv, rowvec_ld=N, rowvec_idx=idx, rows_per_group=rpg, K=K, c_off=64)
        ref = ref + rv[idx.long()].repeat_interleave(rpg, 0)[:M]
    out = E.linear(ctx, x, w, N, **kw)
    with L.knob("NO_TC3"):
        one = E.linear(ctx, x, w, N, **kw)
    assert out.dtype == torch.bfloat16 and rel_err(out, ref) < 6e-3
    assert torch.equal(out, one)
