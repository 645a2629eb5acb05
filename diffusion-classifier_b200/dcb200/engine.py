"""Host-side op layer: thin Python wrappers that turn torch device buffers into C-ABI calls (include/dcb200.h).

Everything here is plumbing -- buffer allocation (torch caching allocator), pointer extraction, stream handoff.
All arithmetic happens in libdcb200.so.  Activations are NHWC in ``ctx.tdtype`` (bf16 fast path / fp32 verify).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import torch

from . import _lib as L


@dataclass
class Ctx:
    device: torch.device
    precision: str = "bf16"      # "bf16" (tcgen05 engine) | "fp32" (CUDA-core verify engine)
    engine: int = L.ENGINE_AUTO  # force an engine for GEMMs (tests: SIMT on bf16 data)

    @property
    def code(self):
        return L.BF16 if self.precision == "bf16" else L.F32

    @property
    def tdtype(self):
        return torch.bfloat16 if self.precision == "bf16" else torch.float32

    def stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def empty(self, *shape, dtype=None):
        return torch.empty(*shape, device=self.device, dtype=dtype or self.tdtype)


class GemmProfile:
    """bench.py's roofline probe: CUDA-event pairs around every tcgen05 GEMM launch + its algorithmic FLOPs."""

    def __init__(self):
        self.rows = []     # (start_event, end_event, flops, tag)
        self.gn_rows = []  # (start_event, end_event, algorithmic bytes) of every streaming GroupNorm-apply launch
        self.attn_rows = []  # (start_event, end_event, 4 B h N^2 d) of every tcgen05 attention call (N >= 128)

    def attn_totals(self):
        torch.cuda.synchronize()
        return (sum(a.elapsed_time(b) for a, b, _ in self.attn_rows), sum(f for _, _, f in self.attn_rows),
                len(self.attn_rows))

    def gn_totals(self):
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b, _ in self.gn_rows), sum(f for _, _, f in self.gn_rows), len(self.gn_rows)

    def totals(self):
        torch.cuda.synchronize()
        ms = sum(a.elapsed_time(b) for a, b, _, _ in self.rows)
        return ms, sum(f for _, _, f, _ in self.rows), len(self.rows)

    def by_shape(self):
        torch.cuda.synchronize()
        out = {}
        for a, b, f, tag in self.rows:
            r = out.setdefault(tag, [0, 0.0, 0.0])
            r[0] += 1
            r[1] += a.elapsed_time(b)
            r[2] += f
        return out


PROFILE = None
PACK_GEN = __import__("itertools").count(1)   # generation counter of packed weight sets (UNetCondition2D / DiT .packed())
USE_TILE_STATS = __import__("os").environ.get("DCB_TILE_STATS", "1") != "0"
USE_FUSED_SMALL_GN = __import__("os").environ.get("DCB_FUSED_SMALL_GN", "1") != "0"
FOLD_UPSAMPLE = __import__("os").environ.get("DCB_FOLD_UPSAMPLE", "1") != "0"   # A/B switch for upsample_conv
FUSE_GN = __import__("os").environ.get("DCB_FUSE_GN", "1") != "0"   # A/B switch: GroupNorm applied inside the consumer conv
# ... only for convs over at least this many input channels (A/B knob; 0 = every eligible conv, the measured best:
# profiles/r02_gn_fusion.md)
FUSE_GN_MIN_C = int(__import__("os").environ.get("DCB_FUSE_GN_MIN_C", "0"))
# conv_out (N = out_channels, fused eps-MSE epilogue) has almost no tensor work to hide the transform behind: the fused launch
# is MUFU bound and measured 0.93-1.0 ms against 0.73-0.83 ms for gn_apply + conv (200 samples of 128^2): not fused
FUSE_GN_MSE = __import__("os").environ.get("DCB_FUSE_GN_MSE", "0") != "0"
XF_UNSUPPORTED = object()     # gemm(xf=...) sentinel: this launch cannot apply the fused transform; nothing was launched


def params_version(module):
    """(version counter, address) of every parameter of ``module``, walked iteratively over ``_modules`` / ``_parameters``
    (0.2 ms for the 491 tensors of unet-128 instead of the 2 ms two ``module.parameters()`` generators take: this runs at
    the top of every classify() call, ahead of the first launch).  Catches in-place updates (optimizer, load_state_dict,
    EMA copies), ``.to()`` moves and replaced parameters / sub-modules alike."""
    out, stack = [], [module]
    while stack:
        m = stack.pop()
        for t in m._parameters.values():
            if t is not None:
                out.append(t._version)
                out.append(t.data_ptr())
        for c in m._modules.values():
            if c is not None:
                stack.append(c)
    return tuple(out)


def _p(t):
    return None if t is None else t.data_ptr()


def seg(src, C_, H, W, c_off=0, kc=None, dy=0, dx=0, stride=1, nb_div=1):
    """nb_div > 1: ``src`` holds one tensor per (image, timestep) unit and output sample n reads unit n // nb_div."""
    return (src, C_, H, W, c_off, C_ - c_off if kc is None else kc, dy, dx, stride, nb_div)


def conv3x3_segs(src, C_, H, W, stride=1, nb_div=1):
    return [seg(src, C_, H, W, 0, C_, ky - 1, kx - 1, stride, nb_div) for ky in range(3) for kx in range(3)]


def gemm(ctx: Ctx, segs, W, N, NB, OH, OW, *, bias=None, rowvec=None, rowvec_ld=0, rowvec_idx=None, rows_per_group=0,
         gate=None, gate_ld=0, residual=None, res_ld=0, res_mod=0, res_idx=None, act=L.ACT_NONE, act_post=L.ACT_NONE,
         out=None, out_dtype=None, out_ld=None, mse=None, want_out=True, k_alg=None, gn_stats=False, up_phase=0,
         gn_part=None, xf=None, attn_norms=None):
    """D = sum_seg A_seg . W^T with the fused epilogue; returns the [M, n_out] output (or None if want_out=False).

    mse = dict(target=, scale=, div=, ld=, err=[S] fp32 out) enables the fused eps-MSE epilogue (tcgen05 only).
    gn_stats=True returns ``(out, part)``: part [ceil(M/128), n_out, 2] fp32 holds per-128-row-tile, per-column
    (sum, sumsq) of the written values -- the next GroupNorm's statistics -- or None when this launch cannot
    produce them (fp32 verify engine, 256-wide direct-epilogue tiles).
    up_phase = 1 + 2a + b: this launch is phase (a, b) of a folded 2x nearest upsample + conv (see ``upsample_conv``);
    ``out`` ([NB*4*OH*OW, n_out]) and ``gn_part`` are then shared by the four phase launches and supplied by the caller.
    xf = dict(a=, b=, src1=, c1=, div1=, silu=, prepare=): segments 0..8 name the RAW input of a GroupNorm(+SiLU) and the
    kernel normalises the operand on the fly (dcb_gemm_desc.xf_a).  Returns ``XF_UNSUPPORTED`` -- before anything is
    launched -- when this launch cannot do that; otherwise calls ``xf['prepare']()`` (fills the coefficient tables) first.
    attn_norms = (ws, heads, tokens): this projection writes q and k of an attention layer (head dim 64); the launch also
    leaves max |q_i|^2, max |k_j|^2 per (sample, head) in ``ws`` (dcb_gemm_desc.attn_norms) for ``attention(norms_ready=ws)``.
    """
    lib = L.lib()
    d = L.GemmDesc()
    d.dtype, d.engine = ctx.code, ctx.engine
    d.NB, d.OH, d.OW, d.N = NB, OH, OW, N
    d.nseg = len(segs)
    keep = []
    for i, (src, C_, H, W_, c_off, kc, dy, dx, stride, nb_div) in enumerate(segs):
        s = d.seg[i]
        s.src, s.C, s.H, s.W, s.c_off, s.kc, s.dy, s.dx, s.stride = src.data_ptr(), C_, H, W_, c_off, kc, dy, dx, stride
        s.nb_div = nb_div
        keep.append(src)
    M = NB * OH * OW
    n_out = N // 2 if act == L.ACT_GEGLU else N
    d.W, d.bias = W.data_ptr(), _p(bias)
    d.rowvec, d.rowvec_ld, d.rowvec_idx, d.rows_per_group = _p(rowvec), rowvec_ld, _p(rowvec_idx), rows_per_group
    d.gate, d.gate_ld = _p(gate), gate_ld
    d.residual, d.res_idx, d.res_ld, d.res_mod = _p(residual), _p(res_idx), res_ld, res_mod
    d.res_dtype = L.F32 if (residual is not None and residual.dtype == torch.float32) else L.BF16
    d.act, d.act_post = act, act_post
    d.up_phase = up_phase
    if attn_norms is not None:
        d.attn_norms, d.attn_heads, d.attn_tok = attn_norms[0].data_ptr(), attn_norms[1], attn_norms[2]
    assert up_phase == 0 or out is not None
    if xf is not None:
        d.xf_a, d.xf_b, d.xf_silu = xf["a"].data_ptr(), xf["b"].data_ptr(), int(xf["silu"])
        d.xf_src1, d.xf_c1, d.xf_div1 = _p(xf.get("src1")), xf.get("c1", 0), xf.get("div1", 1)
        keep.append(xf.get("src1"))
    odt = out_dtype or ctx.tdtype
    if want_out and out is None:
        out = torch.empty(M, n_out if out_ld is None else out_ld, device=ctx.device, dtype=odt)
    if out is not None:
        d.out = out.data_ptr()
        d.out_ld = out_ld if out_ld is not None else (out.shape[-1] if out.dim() == 2 else n_out)
        d.out_dtype = L.F32 if out.dtype == torch.float32 else L.BF16
    gpart = None
    if gn_stats and ctx.code == L.BF16 and out is not None:
        # Ask for the layer shape at a canonical LARGE batch: the tile width (hence staged vs direct epilogue, hence whether
        # tile statistics exist) shrinks with the tile count, and a small batch must not take a different statistics path
        # than a large one -- a sample's result has to be independent of the batch it is scored in.  (Staged at the large
        # batch implies staged at every smaller one: tiles only get narrower.)
        ok = C.c_int32()
        d.NB = max(NB, 1 << 14)
        L.check(lib.dcb_gemm_gn_layout(C.byref(d), C.byref(ok)), "gemm_gn_layout")
        d.NB = NB
        if ok.value:
            gpart = gn_part if gn_part is not None else \
                torch.empty((M + 127) // 128, n_out, 2, device=ctx.device, dtype=torch.float32)
            d.gn_part = gpart.data_ptr()
    part = None
    if mse is not None:
        d.mse_target, d.mse_scale = mse["target"].data_ptr(), _p(mse.get("scale"))
        d.mse_div, d.mse_ld = mse.get("div", 1), mse["ld"]
        rpp, nt = C.c_int32(), C.c_int32()
        d.mse_part = 1  # placeholder so the descriptor validates as "produces something"
        L.check(lib.dcb_gemm_mse_layout(C.byref(d), C.byref(rpp), C.byref(nt)), "gemm_mse_layout")
        assert (OH * OW) % rpp.value == 0
        pps = (OH * OW) // rpp.value * nt.value
        part = torch.empty(NB * pps, device=ctx.device, dtype=torch.float32)
        d.mse_part = part.data_ptr()
    if xf is not None:
        ok = C.c_int32()
        L.check(lib.dcb_gemm_xf_layout(C.byref(d), C.byref(ok)), "gemm_xf_layout")
        if not ok.value:
            return XF_UNSUPPORTED
        xf["prepare"]()
    prof = PROFILE if (PROFILE is not None and ctx.code == L.BF16 and ctx.engine != L.ENGINE_SIMT) else None
    if prof is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    L.check(lib.dcb_gemm(C.byref(d), ctx.stream()), "gemm")
    if prof is not None:
        e1.record()
        K = sum(sg[5] for sg in segs) + (9 * xf.get("c1", 0) if xf is not None else 0)
        prof.rows.append((e0, e1, 2.0 * M * N * (k_alg or K), f"M{M}_N{N}_K{K}_seg{len(segs)}"))
    if mse is not None:
        L.check(lib.dcb_mse_finalize(part.data_ptr(), pps, NB, mse["err"].data_ptr(), 1, ctx.stream()), "mse_finalize")
    if gn_stats:
        return out, gpart
    return out


def linear(ctx, x, W, N, *, K=None, c_off=0, **kw):
    """x: [M, C] row-major tokens; uses channels [c_off, c_off+K)."""
    M, C_ = x.shape
    return gemm(ctx, [seg(x, C_, 1, M, c_off, K if K is not None else C_ - c_off)], W, N, 1, 1, M, **kw)


def upsample_conv(ctx, x, wph, bias, C_, N, NB, H, W):
    """Upsample2D (nearest 2x, then conv3x3 pad 1) without the upsampled tensor: output pixel (2y + a, 2x + b) only ever
    sees a 2x2 neighbourhood of the LOW-resolution input (rows y - 1 + a, y + a; columns x - 1 + b, x + b), with the 3x3
    taps that land on the same source pixel summed -- four 2x2-tap convs (K = 4 C instead of 9 C: 2.25x fewer FLOPs,
    no 4x larger intermediate).  wph[2a + b]: [N, (ty, tx, c)] phase weights (``fold_upsample_weights``).
    x: [NB*H*W, C_] -> ([NB*2H*2W, N], GroupNorm tile statistics or None)."""
    out = ctx.empty(NB * 4 * H * W, N)
    want = H * W % 128 == 0                     # tile statistics need tiles that lie inside one sample
    gp = torch.empty(NB * 4 * H * W // 128, N, 2, device=ctx.device, dtype=torch.float32) if want else None
    st = None
    for a in range(2):
        for b in range(2):
            segs = [seg(x, C_, H, W, 0, C_, a - 1 + ty, b - 1 + tx) for ty in range(2) for tx in range(2)]
            r = gemm(ctx, segs, wph[2 * a + b], N, NB, H, W, bias=bias, out=out, up_phase=1 + 2 * a + b,
                     gn_stats=want, gn_part=gp)
            st = r[1] if want else None
    return out, st


def fold_upsample_weights(w):
    """[Cout, Cin, 3, 3] conv weight (fp32) -> four [Cout, (ty, tx, Cin)] phase weights of ``upsample_conv``:
    phase a = 0 reads source rows (y - 1, y) with taps (ky0, ky1 + ky2); a = 1 reads (y, y + 1) with (ky0 + ky1, ky2)."""
    grp = {0: ((0,), (1, 2)), 1: ((0, 1), (2,))}
    out = []
    for a in range(2):
        for b in range(2):
            taps = []
            for ty in range(2):
                for tx in range(2):
                    acc = 0
                    for ky in grp[a][ty]:
                        for kx in grp[b][tx]:
                            acc = acc + w[:, :, ky, kx]
                    taps.append(acc)
            out.append(torch.stack(taps, 1).reshape(w.shape[0], -1).contiguous())
    return out


def gn_chunks(NB, HW, Ctot):
    """pixel chunks per sample for the GroupNorm statistics pass.  Deliberately a function of the per-sample shape
    only (never of the batch), so a sample's result is bit-identical however the batch is composed / sharded."""
    return max(1, min((HW * Ctot) // 131072, 64, HW))


def groupnorm(ctx, x0, C0, x1, C1, NB, HW, gamma, beta, eps, silu, G=32, div0=1, div1=1, st0=None, st1=None):
    """GroupNorm(+SiLU) over cat([x0, x1], channel) for NB samples; sample n reads x0[n // div0], x1[n // div1].
    st0 / st1: tile statistics written by the GEMMs that produced x0 / x1 (``gemm(gn_stats=True)``); when every source
    has them the statistics pass -- a full re-read of the tensors -- is replaced by a reduction of those partials."""
    lib = L.lib()
    out = ctx.empty(NB * HW, C0 + C1)
    if st0 is not None and (x1 is None or st1 is not None) and HW % 128 == 0 and USE_TILE_STATS:
        chunks = 1
        part = torch.empty(NB * G * 2, device=ctx.device, dtype=torch.float32)
        L.check(lib.dcb_groupnorm_stats_from_tiles(st0.data_ptr(), C0, div0, _p(st1), C1, div1, NB, HW // 128, G,
                                                   part.data_ptr(), ctx.stream()), "groupnorm_stats_from_tiles")
    elif USE_FUSED_SMALL_GN and HW < 128 and ((C0 + C1) // G) % (8 if ctx.code == L.BF16 else 4) == 0 \
            and C0 % ((C0 + C1) // G) == 0 and (C0 + C1) <= 256 * (8 if ctx.code == L.BF16 else 4) \
            and HW * (C0 + C1) <= 32768:
        # (one block walks the whole sample twice: beyond ~32 k elements -- the 8^2 x 1024-channel level of unet-128 -- the
        #  two-kernel path with several blocks per sample is faster: 15.7 vs 11.8 ms of gn_apply per 200 evals, ncu)
        # small samples below a GEMM tile: statistics + apply in one launch, one block per sample (a function of the
        # per-sample shape only, so results stay independent of the batch composition)
        L.check(lib.dcb_groupnorm_fused(ctx.code, x0.data_ptr(), C0, div0, _p(x1), C1, div1, NB, HW, G, gamma.data_ptr(),
                                        beta.data_ptr(), eps, int(silu), out.data_ptr(), ctx.stream()), "groupnorm_fused")
        return out
    else:
        chunks = gn_chunks(NB, HW, C0 + C1)
        part = torch.empty(NB * chunks * G * 2, device=ctx.device, dtype=torch.float32)
        L.check(lib.dcb_groupnorm_stats_div(ctx.code, x0.data_ptr(), C0, div0, _p(x1), C1, div1, NB, HW, G, chunks,
                                            part.data_ptr(), ctx.stream()), "groupnorm_stats")
    prof = PROFILE if (PROFILE is not None and ctx.code == L.BF16) else None
    if prof is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    L.check(lib.dcb_groupnorm_apply_div(ctx.code, x0.data_ptr(), C0, div0, _p(x1), C1, div1, NB, HW, G, chunks,
                                        part.data_ptr(), gamma.data_ptr(), beta.data_ptr(), eps, int(silu),
                                        out.data_ptr(), ctx.stream()), "groupnorm_apply")
    if prof is not None:
        e1.record()     # algorithmic bytes: one read + one write of the normalised tensor
        prof.gn_rows.append((e0, e1, 2.0 * NB * HW * (C0 + C1) * out.element_size()))
    return out


def gn_conv3x3(ctx, x0, C0, x1, C1, NB, H, W, gamma, beta, eps, silu, Wt, N, *, div1=1, st0=None, st1=None, G=32,
               extra_segs=(), **kw):
    """conv3x3(GroupNorm(+SiLU)(cat([x0, x1], channel))) [+ extra K segments, e.g. a 1x1 shortcut over raw tensors].

    Fused form (full-resolution layers: rows of >= 128 pixels, enough tiles for the 256-pixel CTA kernel, producer-written
    tile statistics): ``dcb_groupnorm_coef_from_tiles`` turns the statistics into per-(sample, channel) coefficients and the
    conv normalises its operand on the fly -- the normalised tensor (one HBM write + one read of every activation, the
    ``gn_apply`` pass) never exists.  Bit-identical to the unfused form, which everything else takes."""
    HW = H * W
    if FUSE_GN and ctx.code == L.BF16 and ctx.engine != L.ENGINE_SIMT and st0 is not None and (x1 is None or st1 is not None) \
            and USE_TILE_STATS and HW % 128 == 0 and W % 128 == 0 and C0 % 64 == 0 and C1 % 64 == 0 \
            and C0 + C1 >= FUSE_GN_MIN_C and (FUSE_GN_MSE or kw.get("mse") is None):
        ca = torch.empty(NB, C0 + C1, device=ctx.device, dtype=torch.float32)
        cb = torch.empty(NB, C0 + C1, device=ctx.device, dtype=torch.float32)

        def prepare():
            L.check(L.lib().dcb_groupnorm_coef_from_tiles(st0.data_ptr(), C0, 1, _p(st1), C1, div1, NB, HW // 128, G,
                                                          gamma.data_ptr(), beta.data_ptr(), eps, ca.data_ptr(),
                                                          cb.data_ptr(), ctx.stream()), "groupnorm_coef_from_tiles")

        r = gemm(ctx, conv3x3_segs(x0, C0, H, W) + list(extra_segs), Wt, N, NB, H, W,
                 xf=dict(a=ca, b=cb, src1=x1, c1=C1, div1=div1, silu=silu, prepare=prepare), **kw)
        if r is not XF_UNSUPPORTED:
            return r
    a = groupnorm(ctx, x0, C0, x1, C1, NB, HW, gamma, beta, eps, silu, G=G, div1=div1, st0=st0, st1=st1)
    return gemm(ctx, conv3x3_segs(a, C0 + C1, H, W) + list(extra_segs), Wt, N, NB, H, W, **kw)


def expand_samples(ctx, x, NB, div, rows_per_sample):
    """materialised class expansion [NB/div * rows, C] -> [NB * rows, C] (small low-resolution tensors only)."""
    C_ = x.shape[1]
    out = torch.empty(NB * rows_per_sample, C_, device=ctx.device, dtype=x.dtype)
    L.check(L.lib().dcb_expand_samples(L.F32 if x.dtype == torch.float32 else L.BF16, x.data_ptr(), NB, div,
                                       rows_per_sample * C_, out.data_ptr(), ctx.stream()), "expand_samples")
    return out


def layernorm(ctx, x, gamma=None, beta=None, eps=1e-5, scale=None, shift=None, mod_ld=0, rows_per_group=0):
    rows, C_ = x.shape
    out = torch.empty_like(x)
    L.check(L.lib().dcb_layernorm(ctx.code, x.data_ptr(), rows, C_, _p(gamma), _p(beta), eps, _p(scale), _p(shift),
                                  mod_ld, rows_per_group, out.data_ptr(), ctx.stream()), "layernorm")
    return out


def attn_norms_ws(ctx, B, heads):
    """workspace of the attention pre-pass: [2 + 2 B heads] fp32 (max |q|^2, max |k|^2 per (sample, head))"""
    return torch.empty(2 + B * heads * 2, device=ctx.device, dtype=torch.float32)


def attention(ctx, qkv, B, Ntok, heads, d, q_off=0, k_off=None, v_off=None, simt=False, scale=None, norms_ready=None):
    """qkv: [B*Ntok, ld] with q/k/v column blocks of width heads*d starting at q_off/k_off/v_off.
    scale: softmax scale (default d^-0.5); a caller that folded d^-0.5 log2(e) into its query projection passes 1 / log2(e).
    norms_ready: the ``attn_norms_ws`` buffer the projection that wrote ``qkv`` filled (``gemm(attn_norms=)``): the
    launch skips its own pass over q and k."""
    Cw = heads * d
    ld = qkv.shape[1]
    k_off = Cw if k_off is None else k_off
    v_off = 2 * Cw if v_off is None else v_off
    es = qkv.element_size()
    base = qkv.data_ptr()
    out = torch.empty(B * Ntok, Cw, device=ctx.device, dtype=qkv.dtype)
    code = ctx.code | (0x100 if (simt and ctx.code == L.BF16) else 0)
    ws = attn_norms_ws(ctx, B, heads) if code == L.BF16 and d == 64 else None
    if norms_ready is not None and code == L.BF16 and d == 64:
        ws, code = norms_ready, code | L.ATTN_NORMS_READY
    prof = PROFILE if (PROFILE is not None and (code & 0xFF) == L.BF16 and not simt and Ntok >= 128) else None
    if prof is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    L.check(L.lib().dcb_attention_ws(code, base + q_off * es, base + k_off * es, base + v_off * es, ld, B, Ntok, heads, d,
                                     float(d) ** -0.5 if scale is None else float(scale), out.data_ptr(), Cw, _p(ws),
                                     ctx.stream()), "attention")
    if prof is not None:
        e1.record()
        prof.attn_rows.append((e0, e1, 4.0 * B * heads * Ntok * Ntok * d))
    return out


def upsample2x(ctx, x, NB, H, W, C_):
    out = ctx.empty(NB * 4 * H * W, C_)
    L.check(L.lib().dcb_upsample2x(ctx.code, x.data_ptr(), NB, H, W, C_, out.data_ptr(), ctx.stream()), "upsample2x")
    return out


def timestep_embed(ctx, t, U, rep, dim, shift, max_period=10000.0):
    out = ctx.empty(U * rep, dim)
    L.check(L.lib().dcb_timestep_embed(ctx.code, t.data_ptr(), U, rep, dim, float(shift), float(max_period),
                                       out.data_ptr(), ctx.stream()), "timestep_embed")
    return out


def prologue(ctx, mode, x, U, rep, C_, H, W, kpad, *, patch=1, eps=None, seed=0, unit_id0=0, alpha=None, sigma=None,
             img=None, want_target=False, v_param=False, a_out=None, target_out=None):
    """q_sample + first-layer operand staging.  Returns (a_in [U*rep*rows, kpad], target or None).
    ``a_out`` / ``target_out`` let the caller supply persistent buffers (CUDA-graph replay reads fixed addresses)."""
    rows = H * W if mode == 0 else (H // patch) * (W // patch)
    z_ws = torch.empty(U * H * W * C_, device=ctx.device, dtype=torch.float32)
    a = a_out if a_out is not None else ctx.empty(U * rep * rows, kpad)
    tgt = target_out if target_out is not None else (
        torch.empty(U * H * W * C_, device=ctx.device, dtype=torch.float32) if want_target else None)
    L.check(L.lib().dcb_prologue(mode, ctx.code, x.data_ptr(), _p(eps), seed, unit_id0, _p(alpha), _p(sigma), _p(img), U,
                                 rep, C_, H, W, patch, kpad, z_ws.data_ptr(), a.data_ptr(), _p(tgt), int(v_param),
                                 ctx.stream()), "prologue")
    return a, tgt


def ddpm_step(ctx, z_t, pred, rep, patch, coef, v_param, final, noise=None, seed=0, unit_id0=0, out=None):
    """one ancestral DDPM step + classifier-free guidance (dcb_ddpm_step).  z_t [B,C,H,W] fp32; pred: fp32 output of the
    denoiser's last GEMM (sample b*rep conditional, b*rep+1 unconditional); coef: device [8] fp32."""
    B, C_, H, W = z_t.shape
    assert z_t.dtype == torch.float32 and pred.dtype == torch.float32 and z_t.is_contiguous() and pred.is_contiguous()
    out = torch.empty_like(z_t) if out is None else out
    L.check(L.lib().dcb_ddpm_step(z_t.data_ptr(), pred.data_ptr(), rep, patch, coef.data_ptr(), int(v_param), int(final),
                                  _p(noise), seed, unit_id0, B, C_, H, W, out.data_ptr(), ctx.stream()), "ddpm_step")
    return out


def eps_mse(ctx, pred, target, scale, S, div, K, err):
    L.check(L.lib().dcb_eps_mse(L.F32 if pred.dtype == torch.float32 else L.BF16, pred.data_ptr(), target.data_ptr(),
                                _p(scale), S, div, K, err.data_ptr(), 1, ctx.stream()), "eps_mse")


def nhwc_to_nchw(ctx, x, NB, HW, C_, ld):
    out = torch.empty(NB, C_, HW, device=ctx.device, dtype=torch.float32)
    L.check(L.lib().dcb_nhwc_to_nchw(L.F32 if x.dtype == torch.float32 else L.BF16, x.data_ptr(), NB, HW, C_, ld,
                                     out.data_ptr(), ctx.stream()), "nhwc_to_nchw")
    return out


def unpatchify(ctx, tok, B, g, p, C_, ld):
    out = torch.empty(B, C_, g * p, g * p, device=ctx.device, dtype=torch.float32)
    L.check(L.lib().dcb_unpatchify(L.F32 if tok.dtype == torch.float32 else L.BF16, tok.data_ptr(), B, g, p, C_, ld,
                                   out.data_ptr(), ctx.stream()), "unpatchify")
    return out


def cast(ctx, src_f32, tdtype=None):
    """fp32 parameter -> engine dtype (bf16 rounding done by our kernel, same as everywhere else)."""
    tdtype = tdtype or ctx.tdtype
    src = src_f32.detach().to(device=ctx.device, dtype=torch.float32).contiguous()
    if tdtype == torch.float32:
        return src
    dst = torch.empty(src.shape, device=ctx.device, dtype=tdtype)
    L.check(L.lib().dcb_cast_f32(L.BF16, src.data_ptr(), src.numel(), dst.data_ptr(), ctx.stream()), "cast_f32")
    return dst


# ---- weight packing through the C ABI (dcb_pack_*): checkpoint layout (fp32) -> GEMM operand layouts, on the device ---------
def _f32dev(ctx, t):
    return t.detach().to(device=ctx.device, dtype=torch.float32).contiguous()


def pack_conv(ctx, w, kpad=None, out=None):
    """conv weight [Cout, Cin, kh, kw] -> [Cout, kpad] with K order (ky, kx, cin) (kpad >= kh*kw*Cin: zero padded)."""
    Cout, Cin, kh, kw = w.shape
    kpad = kpad or kh * kw * Cin
    out = ctx.empty(Cout, kpad) if out is None else out
    src = _f32dev(ctx, w)
    L.check(L.lib().dcb_pack_conv(ctx.code, src.data_ptr(), Cout, Cin, kh, kw, kpad, out.data_ptr(), ctx.stream()), "pack_conv")
    return out


def pack_rows(ctx, mats, axis=0, out=None, col0=0):
    """concatenate fp32 [rows, cols] matrices along rows (axis 0) or columns (axis 1) into one operand in the engine dtype;
    ``out`` / ``col0``: write into an existing packed matrix starting at that column (axis 1 only)."""
    mats = [_f32dev(ctx, m.reshape(m.shape[0], -1)) for m in mats]
    if axis == 0:
        rows, cols = sum(m.shape[0] for m in mats), mats[0].shape[1]
        out = ctx.empty(rows, cols) if out is None else out
        r0 = 0
        for m in mats:
            L.check(L.lib().dcb_pack_rows(ctx.code, m.data_ptr(), m.shape[1], m.shape[0], m.shape[1], out.data_ptr(),
                                          out.shape[1], r0, col0, ctx.stream()), "pack_rows")
            r0 += m.shape[0]
    else:
        rows = mats[0].shape[0]
        out = ctx.empty(rows, sum(m.shape[1] for m in mats)) if out is None else out
        for m in mats:
            L.check(L.lib().dcb_pack_rows(ctx.code, m.data_ptr(), m.shape[1], rows, m.shape[1], out.data_ptr(),
                                          out.shape[1], 0, col0, ctx.stream()), "pack_rows")
            col0 += m.shape[1]
    return out


def pack_geglu(ctx, w, b):
    """diffusers GEGLU proj [2*inner, C] (+ bias) -> (rows interleaved per 128 [value | gate], fp32 bias in the same order)."""
    inner, C_ = w.shape[0] // 2, w.shape[1]
    wo = ctx.empty(2 * inner, C_)
    bo = torch.empty(2 * inner, device=ctx.device, dtype=torch.float32)
    ws, bs = _f32dev(ctx, w), _f32dev(ctx, b)
    L.check(L.lib().dcb_pack_geglu(ctx.code, ws.data_ptr(), bs.data_ptr(), inner, C_, wo.data_ptr(), bo.data_ptr(),
                                   ctx.stream()), "pack_geglu")
    return wo, bo


def pack_upsample(ctx, w):
    """[Cout, Cin, 3, 3] -> the four [Cout, (ty, tx, Cin)] phase weights of ``upsample_conv`` (== fold_upsample_weights)."""
    Cout, Cin = w.shape[:2]
    out = ctx.empty(4, Cout, 4 * Cin)
    src = _f32dev(ctx, w)
    L.check(L.lib().dcb_pack_upsample(ctx.code, src.data_ptr(), Cout, Cin, out.data_ptr(), ctx.stream()), "pack_upsample")
    return [out[i] for i in range(4)]


def haar_dwt(x, post_scale=1.0):
    B, C_, H, W = x.shape
    x = x.contiguous().float()
    out = torch.empty(B, 4 * C_, H // 2, W // 2, device=x.device, dtype=torch.float32)
    L.check(L.lib().dcb_haar_dwt(x.data_ptr(), B, C_, H, W, float(post_scale), out.data_ptr(),
                                 torch.cuda.current_stream(x.device).cuda_stream), "haar_dwt")
    return out


def haar_idwt(w, pre_scale=1.0):
    B, C4, h, wd = w.shape
    w = w.contiguous().float()
    out = torch.empty(B, C4 // 4, 2 * h, 2 * wd, device=w.device, dtype=torch.float32)
    L.check(L.lib().dcb_haar_idwt(w.data_ptr(), B, C4, h, wd, float(pre_scale), out.data_ptr(),
                                  torch.cuda.current_stream(w.device).cuda_stream), "haar_idwt")
    return out
