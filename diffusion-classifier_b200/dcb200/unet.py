"""UNetCondition2D -- drop-in for the reference's nets/unet.py:77-195 (a kwargs passthrough to
diffusers.UNet2DConditionModel 0.31.0), executed by libdcb200's sm_100a kernels.

Same constructor kwargs, same ``forward(x, noise_labels, downblock_additional_residuals=None,
midblock_additional_residuals=None, encoder_hidden_states=None)`` and the same state_dict keys as diffusers
 so reference checkpoints load unchanged.  The nn.Modules below are parameter
containers only: no torch arithmetic runs on the hot path (and none exists as a fallback).

Exact algebra used (SURVEY.md finding 3): the class/text context is a single KV token, so softmax over one
key is exactly 1 and attn2 == to_out(to_v(ctx)) -- a per-sample [C] vector, added in the attn1 out-proj epilogue.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Optional, Tuple, Union

import torch
import torch.nn as nn

from . import _lib as L
from . import engine as E


# ---- parameter containers (names == diffusers state_dict keys) ---------------------------------------------
class _P(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter container; the compute path is libdcb200")


class _Resnet(_P):
    def __init__(self, cin, cout, tdim, groups, eps):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps)
        self.conv1 = nn.Conv2d(cin, cout, 3, 1, 1)
        self.time_emb_proj = nn.Linear(tdim, cout)
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps)
        self.conv2 = nn.Conv2d(cout, cout, 3, 1, 1)
        if cin != cout:
            self.conv_shortcut = nn.Conv2d(cin, cout, 1)
        self.cin, self.cout, self.eps = cin, cout, eps


class _Attn(_P):
    def __init__(self, qdim, ctx_dim, bias=False):
        super().__init__()
        self.to_q = nn.Linear(qdim, qdim, bias=bias)
        self.to_k = nn.Linear(ctx_dim, qdim, bias=bias)
        self.to_v = nn.Linear(ctx_dim, qdim, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(qdim, qdim), nn.Dropout(0.0)])


class _GEGLU(_P):
    def __init__(self, d, inner):
        super().__init__()
        self.proj = nn.Linear(d, inner * 2)


class _FF(_P):
    def __init__(self, d):
        super().__init__()
        self.net = nn.ModuleList([_GEGLU(d, 4 * d), nn.Dropout(0.0), nn.Linear(4 * d, d)])


class _TBlock(_P):
    def __init__(self, d, xdim):
        super().__init__()
        self.norm1 = nn.LayerNorm(d, eps=1e-5)
        self.attn1 = _Attn(d, d)
        self.norm2 = nn.LayerNorm(d, eps=1e-5)
        self.attn2 = _Attn(d, xdim)
        self.norm3 = nn.LayerNorm(d, eps=1e-5)
        self.ff = _FF(d)


class _Transformer2D(_P):
    def __init__(self, ch, heads, xdim, groups):
        super().__init__()
        self.norm = nn.GroupNorm(groups, ch, eps=1e-6)
        self.proj_in = nn.Conv2d(ch, ch, 1)
        self.transformer_blocks = nn.ModuleList([_TBlock(ch, xdim)])
        self.proj_out = nn.Conv2d(ch, ch, 1)
        self.ch, self.heads = ch, heads


class _Sampler(_P):
    def __init__(self, ch, stride):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 3, stride=stride, padding=1)


class _TimeEmb(_P):
    def __init__(self, cin, tdim):
        super().__init__()
        self.linear_1 = nn.Linear(cin, tdim)
        self.linear_2 = nn.Linear(tdim, tdim)


class _Act:
    """activation handle: tensor [nb*HW, C], channel count, sample divisor (1 = per sample, rep = per unit)."""
    __slots__ = ("t", "C", "div", "st")

    def __init__(self, t, C, div, st=None):
        if isinstance(t, tuple):      # (tensor, GroupNorm tile statistics) from gemm(gn_stats=True)
            t, st = t
        self.t, self.C, self.div, self.st = t, C, div, st


def _tup(v, n):
    return tuple(v) if isinstance(v, (tuple, list)) else (v,) * n


class UNetCondition2D(nn.Module):
    def __init__(
        self,
        sample_size: Optional[int] = None,
        in_channels: int = 4,
        out_channels: int = 4,
        center_input_sample: bool = False,
        flip_sin_to_cos: bool = True,
        freq_shift: int = 0,
        down_block_types: Tuple[str] = ("CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D",
                                        "DownBlock2D"),
        mid_block_type: Optional[str] = "UNetMidBlock2DCrossAttn",
        up_block_types: Tuple[str] = ("UpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D"),
        only_cross_attention: Union[bool, Tuple[bool]] = False,
        block_out_channels: Tuple[int] = (320, 640, 1280, 1280),
        layers_per_block: Union[int, Tuple[int]] = 2,
        downsample_padding: int = 1,
        mid_block_scale_factor: float = 1,
        dropout: float = 0.0,
        act_fn: str = "silu",
        norm_num_groups: Optional[int] = 32,
        norm_eps: float = 1e-5,
        cross_attention_dim: Union[int, Tuple[int]] = 1280,
        transformer_layers_per_block=1,
        reverse_transformer_layers_per_block=None,
        encoder_hid_dim: Optional[int] = None,
        encoder_hid_dim_type: Optional[str] = None,
        attention_head_dim: Union[int, Tuple[int]] = 8,
        num_attention_heads: Optional[Union[int, Tuple[int]]] = None,
        dual_cross_attention: bool = False,
        use_linear_projection: bool = False,
        class_embed_type: Optional[str] = None,
        addition_embed_type: Optional[str] = None,
        addition_time_embed_dim: Optional[int] = None,
        num_class_embeds: Optional[int] = None,
        upcast_attention: bool = False,
        resnet_time_scale_shift: str = "default",
        resnet_skip_time_act: bool = False,
        resnet_out_scale_factor: float = 1.0,
        time_embedding_type: str = "positional",
        time_embedding_dim: Optional[int] = None,
        time_embedding_act_fn: Optional[str] = None,
        timestep_post_act: Optional[str] = None,
        time_cond_proj_dim: Optional[int] = None,
        conv_in_kernel: int = 3,
        conv_out_kernel: int = 3,
        projection_class_embeddings_input_dim: Optional[int] = None,
        attention_type: str = "default",
        class_embeddings_concat: bool = False,
        mid_block_only_cross_attention: Optional[bool] = None,
        cross_attention_norm: Optional[str] = None,
        addition_embed_type_num_heads: int = 64,
    ):
        super().__init__()
        unsupported = dict(
            center_input_sample=center_input_sample, only_cross_attention=only_cross_attention,
            dual_cross_attention=dual_cross_attention, use_linear_projection=use_linear_projection,
            class_embed_type=class_embed_type, addition_embed_type=addition_embed_type,
            num_class_embeds=num_class_embeds, upcast_attention=upcast_attention,
            resnet_skip_time_act=resnet_skip_time_act, time_embedding_dim=time_embedding_dim,
            time_embedding_act_fn=time_embedding_act_fn, timestep_post_act=timestep_post_act,
            time_cond_proj_dim=time_cond_proj_dim, class_embeddings_concat=class_embeddings_concat,
            mid_block_only_cross_attention=mid_block_only_cross_attention, cross_attention_norm=cross_attention_norm,
            reverse_transformer_layers_per_block=reverse_transformer_layers_per_block)
        bad = {k: v for k, v in unsupported.items() if v not in (None, False)}
        if bad or mid_block_type != "UNetMidBlock2DCrossAttn" or act_fn != "silu" or dropout != 0.0 \
                or transformer_layers_per_block != 1 or encoder_hid_dim_type != "text_proj" \
                or resnet_time_scale_shift != "default" or time_embedding_type != "positional" \
                or conv_in_kernel != 3 or conv_out_kernel != 3 or attention_type != "default" \
                or downsample_padding != 1 or resnet_out_scale_factor != 1.0 or mid_block_scale_factor != 1 \
                or not flip_sin_to_cos or isinstance(cross_attention_dim, (tuple, list)):
            raise NotImplementedError(
                f"dcb200.UNetCondition2D implements the code paths the reference's configs select "
                f"(models/*.py, experiments/*); unsupported kwargs: {bad}")
        boc = tuple(block_out_channels)
        n = len(boc)
        lpb = _tup(layers_per_block, n)
        heads = _tup(num_attention_heads or attention_head_dim, n)  # diffusers quirk: attention_head_dim==#heads
        G = norm_num_groups
        self.config = SimpleNamespace(
            sample_size=sample_size, in_channels=in_channels, out_channels=out_channels, block_out_channels=boc,
            layers_per_block=layers_per_block, encoder_hid_dim=encoder_hid_dim, cross_attention_dim=cross_attention_dim,
            down_block_types=tuple(down_block_types), up_block_types=tuple(up_block_types), freq_shift=freq_shift,
            norm_num_groups=G, norm_eps=norm_eps)
        tdim = boc[0] * 4
        xd = cross_attention_dim
        self.conv_in = nn.Conv2d(in_channels, boc[0], 3, padding=1)
        self.time_embedding = _TimeEmb(boc[0], tdim)
        self.encoder_hid_proj = nn.Linear(encoder_hid_dim, xd)
        self.down_blocks = nn.ModuleList()
        out = boc[0]
        for i, typ in enumerate(down_block_types):
            inp, out = out, boc[i]
            blk = _P()
            blk.resnets = nn.ModuleList([_Resnet(inp if j == 0 else out, out, tdim, G, norm_eps) for j in range(lpb[i])])
            if typ == "CrossAttnDownBlock2D":
                blk.attentions = nn.ModuleList([_Transformer2D(out, heads[i], xd, G) for _ in range(lpb[i])])
            elif typ != "DownBlock2D":
                raise NotImplementedError(typ)
            if i != n - 1:
                blk.downsamplers = nn.ModuleList([_Sampler(out, 2)])
            self.down_blocks.append(blk)
        Cm = boc[-1]
        self.mid_block = _P()
        self.mid_block.attentions = nn.ModuleList([_Transformer2D(Cm, heads[-1], xd, G)])
        self.mid_block.resnets = nn.ModuleList([_Resnet(Cm, Cm, tdim, G, norm_eps) for _ in range(2)])
        self.up_blocks = nn.ModuleList()
        rb, rl, rh = boc[::-1], lpb[::-1], heads[::-1]
        out = rb[0]
        for i, typ in enumerate(up_block_types):
            prev, out = out, rb[i]
            inn = rb[min(i + 1, n - 1)]
            Lr = rl[i] + 1
            blk = _P()
            blk.resnets = nn.ModuleList(
                [_Resnet((prev if j == 0 else out) + (inn if j == Lr - 1 else out), out, tdim, G, norm_eps)
                 for j in range(Lr)])
            if typ == "CrossAttnUpBlock2D":
                blk.attentions = nn.ModuleList([_Transformer2D(out, rh[i], xd, G) for _ in range(Lr)])
            elif typ != "UpBlock2D":
                raise NotImplementedError(typ)
            if i != n - 1:
                blk.upsamplers = nn.ModuleList([_Sampler(out, 1)])
            self.up_blocks.append(blk)
        self.conv_norm_out = nn.GroupNorm(G, boc[0], eps=norm_eps)
        self.conv_out = nn.Conv2d(boc[0], out_channels, 3, padding=1)
        self._packs = {}
        self._row_idx = {}
        self.precision = "bf16"  # "fp32" selects the CUDA-core verify engine (north star: 1e-4 mode)

    # ---- weight packing (layout plumbing, once per parameter version) --------------------------------------
    def _resnets(self):
        for blk in self.down_blocks:
            yield from blk.resnets
        yield from self.mid_block.resnets
        for blk in self.up_blocks:
            yield from blk.resnets

    def _transformers(self):
        for blk in list(self.down_blocks) + [self.mid_block] + list(self.up_blocks):
            if hasattr(blk, "attentions"):
                yield from blk.attentions

    def _version(self):
        return E.params_version(self)

    def packed(self, ctx: E.Ctx):
        key = (ctx.precision, str(ctx.device))
        ver = self._version()
        hit = self._packs.get(key)
        if hit is not None and hit.version == ver:
            return hit
        pk = self._pack(ctx)
        pk.version = ver
        pk.gen = next(E.PACK_GEN)      # monotonic: CUDA-graph caches key on it (an id() can be reused after a repack)
        self._packs[key] = pk
        return pk

    @torch.no_grad()
    def _pack(self, ctx):
        def w(t):
            return E.cast(ctx, t)

        def f32(t):
            return t.detach().to(ctx.device, torch.float32).contiguous()

        def conv_w(c, kpad=None):  # [Cout,Cin,kh,kw] -> [Cout, (ky,kx,c)] (+ zero padding): dcb_pack_conv
            return E.pack_conv(ctx, c.weight, kpad)

        # every operand layout below is produced by the library's dcb_pack_* entry points (csrc/pack.cu), so a binder that
        # is not Python can build the same packed weights from the checkpoint tensors
        pk = SimpleNamespace()
        cin = self.conv_in.weight.shape[1]
        pk.kpad_in = (9 * cin + 63) // 64 * 64
        pk.conv_in_w, pk.conv_in_b = conv_w(self.conv_in, pk.kpad_in), f32(self.conv_in.bias)
        te = self.time_embedding
        pk.te1_w, pk.te1_b, pk.te2_w, pk.te2_b = w(te.linear_1.weight), f32(te.linear_1.bias), w(te.linear_2.weight), \
            f32(te.linear_2.bias)
        pk.ehp_w, pk.ehp_b = w(self.encoder_hid_proj.weight), f32(self.encoder_hid_proj.bias)
        # every resnet's time_emb_proj as ONE [sum Cout, tdim] GEMM
        res = list(self._resnets())
        pk.temb_w = E.pack_rows(ctx, [r.time_emb_proj.weight for r in res], 0)
        pk.temb_b = f32(torch.cat([r.time_emb_proj.bias.detach() for r in res], 0))
        off = 0
        pk.res = {}
        for r in res:
            q = SimpleNamespace(temb_off=off, cin=r.cin, cout=r.cout, eps=r.eps)
            off += r.cout
            q.g1, q.b1n, q.g2, q.b2n = f32(r.norm1.weight), f32(r.norm1.bias), f32(r.norm2.weight), f32(r.norm2.bias)
            q.w1, q.b1 = conv_w(r.conv1), f32(r.conv1.bias)
            if hasattr(r, "conv_shortcut"):  # fused: [conv2 | 1x1 shortcut] along K, biases summed
                q.w2 = conv_w(r.conv2, 9 * r.cout + r.cin)
                E.pack_rows(ctx, [r.conv_shortcut.weight], 1, out=q.w2, col0=9 * r.cout)
                q.b2 = f32(r.conv2.bias.detach() + r.conv_shortcut.bias.detach())
                q.shortcut = True
            else:
                q.w2, q.b2, q.shortcut = conv_w(r.conv2), f32(r.conv2.bias), False
            pk.res[id(r)] = q
        pk.temb_total = off
        # collapsed cross-attention: v = to_v(ctx) for all layers in one GEMM, then per-layer to_out
        trs = list(self._transformers())
        pk.xv_w = E.pack_rows(ctx, [t.transformer_blocks[0].attn2.to_v.weight for t in trs], 0)
        off = 0
        pk.tr = {}
        for t in trs:
            b = t.transformer_blocks[0]
            Cc = t.ch
            q = SimpleNamespace(xv_off=off, ch=Cc, heads=t.heads)
            off += Cc
            q.gn_g, q.gn_b = f32(t.norm.weight), f32(t.norm.bias)
            q.pin_w, q.pin_b = w(t.proj_in.weight.detach().reshape(Cc, Cc)), f32(t.proj_in.bias)
            q.pout_w, q.pout_b = w(t.proj_out.weight.detach().reshape(Cc, Cc)), f32(t.proj_out.bias)
            q.ln1_g, q.ln1_b, q.ln3_g, q.ln3_b = f32(b.norm1.weight), f32(b.norm1.bias), f32(b.norm3.weight), \
                f32(b.norm3.bias)
            q.qkv_w = E.pack_rows(ctx, [b.attn1.to_q.weight, b.attn1.to_k.weight, b.attn1.to_v.weight], 0)
            q.o1_w, q.o1_b = w(b.attn1.to_out[0].weight), f32(b.attn1.to_out[0].bias)
            q.xo_w, q.xo_b = w(b.attn2.to_out[0].weight), f32(b.attn2.to_out[0].bias)
            # rows [0,inner) = value, [inner,2*inner) = gate (diffusers GEGLU chunk order) -> interleaved per 128
            q.gg_w, q.gg_b = E.pack_geglu(ctx, b.ff.net[0].proj.weight, b.ff.net[0].proj.bias)
            q.ff2_w, q.ff2_b = w(b.ff.net[2].weight), f32(b.ff.net[2].bias)
            pk.tr[id(t)] = q
        pk.xv_total = off
        pk.samp = {}
        for blk in list(self.down_blocks) + list(self.up_blocks):
            for name in ("downsamplers", "upsamplers"):
                if hasattr(blk, name):
                    s = getattr(blk, name)[0]
                    pk.samp[id(s)] = SimpleNamespace(w=conv_w(s.conv), b=f32(s.conv.bias))
                    if name == "upsamplers":   # nearest-2x + conv3x3 folded into four 2x2-tap phase convs
                        pk.samp[id(s)].wph = E.pack_upsample(ctx, s.conv.weight)
        pk.out_g, pk.out_bn = f32(self.conv_norm_out.weight), f32(self.conv_norm_out.bias)
        pk.out_w, pk.out_b = conv_w(self.conv_out), f32(self.conv_out.bias)
        return pk

    # ---- the denoiser program -------------------------------------------------------------------------------
    # Activations travel as _Act(t, C, div): ``t`` is [nb*HW, C]; div == 1 -> one tensor per sample (nb = S = U*rep),
    # div == rep > 1 -> one tensor per (image, timestep) UNIT (nb = U), shared by the unit's rep class-conditional
    # samples.  Every layer whose inputs are per-unit and that does not see the class (everything up to the first
    # attn1 out-projection, where the collapsed cross-attention vector enters) is computed once per unit; the reference
    # recomputes it for every candidate class (diffusion_classifier.py:694-704 loops classes around the full forward).
    # The arithmetic per sample is unchanged, so results are identical to the unshared program.
    def _unit_rows(self, ctx, S, HW, rep):
        """int32 [S*HW]: row of the per-unit tensor that sample-major row m reads (persistent: CUDA graphs hold it)."""
        key = (S, HW, rep, str(ctx.device))
        idx = self._row_idx.get(key)
        if idx is None:
            m = torch.arange(S * HW, device=ctx.device, dtype=torch.int64)
            idx = ((m // (HW * rep)) * HW + m % HW).to(torch.int32).contiguous()
            self._row_idx[key] = idx
        return idx

    def _resnet(self, ctx, q, temb, x0, x1, U, rep, H, W):
        HW = H * W
        S = U * rep
        unit = x0.div > 1 and (x1 is None or x1.div > 1)
        NB = U if unit else S
        d0 = 1 if unit else x0.div
        assert d0 == 1, "the running activation is per-sample once any class-dependent layer has run"
        d1 = 1 if (unit or x1 is None) else x1.div
        if d1 > 1 and HW < 128:      # a GEMM tile would span samples: materialise the (tiny) expanded skip
            x1 = _Act(E.expand_samples(ctx, x1.t, S, d1, HW), x1.C, 1)
            d1 = 1
        t1, C0 = (x1.t if x1 is not None else None), x0.C
        s1 = x1.st if x1 is not None else None
        C1 = x1.C if x1 is not None else 0
        tld = temb.shape[1] * (rep if unit else 1)   # per-unit layers read the unit's first sample row of temb
        # norm1 + SiLU + conv1 (+ time embedding), norm2 + SiLU + conv2 (+ shortcut / residual): the GroupNorms are applied
        # inside the convs where the launch supports it (E.gn_conv3x3), else as their own streaming pass
        h1, hs = E.gn_conv3x3(ctx, x0.t, C0, t1, C1, NB, H, W, q.g1, q.b1n, q.eps, True, q.w1, q.cout, div1=d1,
                              st0=x0.st, st1=s1, bias=q.b1, rowvec=temb[:, q.temb_off:], rowvec_ld=tld,
                              rows_per_group=HW, gn_stats=True)
        if q.shortcut:
            extra = [E.seg(x0.t, C0, H, W)]
            if x1 is not None:
                extra.append(E.seg(t1, C1, H, W, nb_div=d1))
            out = E.gn_conv3x3(ctx, h1, q.cout, None, 0, NB, H, W, q.g2, q.b2n, q.eps, True, q.w2, q.cout, st0=hs,
                               extra_segs=extra, bias=q.b2, gn_stats=True)
        else:
            out = E.gn_conv3x3(ctx, h1, q.cout, None, 0, NB, H, W, q.g2, q.b2n, q.eps, True, q.w2, q.cout, st0=hs,
                               bias=q.b2, residual=x0.t, res_ld=C0, gn_stats=True)
        return _Act(out, q.cout, rep if unit else 1)

    def _transformer(self, ctx, q, xattn, xattn_idx, x, U, rep, H, W):
        HW, Cc = H * W, q.ch
        S = U * rep
        shared = x.div > 1 and HW >= 128          # attention core once per unit
        NB = U if shared else S
        ridx = self._unit_rows(ctx, S, HW, rep) if x.div > 1 else None
        a = E.groupnorm(ctx, x.t, Cc, None, 0, NB, HW, q.gn_g, q.gn_b, 1e-6, False, div0=1 if shared else x.div, st0=x.st)
        h = E.linear(ctx, a, q.pin_w, Cc, bias=q.pin_b)
        n1 = E.layernorm(ctx, h, q.ln1_g, q.ln1_b, 1e-5)
        qkv = E.linear(ctx, n1, q.qkv_w, 3 * Cc)
        att = E.attention(ctx, qkv, NB, HW, q.heads, Cc // q.heads)
        # attn1 out-proj + residual + collapsed single-token cross-attention (attn2) in one epilogue: the first place a
        # sample sees its class.  With a shared core the A operand / residual rows come from the unit's tensors.
        epi = dict(bias=q.o1_b, residual=h, res_ld=Cc, rowvec=xattn[:, q.xv_off:], rowvec_ld=xattn.shape[1],
                   rowvec_idx=xattn_idx, rows_per_group=HW)
        if shared:
            h = E.gemm(ctx, [E.seg(att, Cc, H, W, nb_div=rep)], q.o1_w, Cc, S, H, W, res_idx=ridx, **epi)
        else:
            h = E.linear(ctx, att, q.o1_w, Cc, **epi)
        n3 = E.layernorm(ctx, h, q.ln3_g, q.ln3_b, 1e-5)
        ff = E.linear(ctx, n3, q.gg_w, 8 * Cc, bias=q.gg_b, act=L.ACT_GEGLU)
        h = E.linear(ctx, ff, q.ff2_w, Cc, bias=q.ff2_b, residual=h, res_ld=Cc)
        assert S * HW == h.shape[0]
        out = E.linear(ctx, h, q.pout_w, Cc, bias=q.pout_b, residual=x.t, res_ld=Cc, res_idx=ridx, gn_stats=True)
        return _Act(out, Cc, 1)

    def cross_attn_table(self, ctx, pk, ehs):
        """ehs [R, hid] (engine dtype) -> fp32 [R, sum C_layer]: attn2 output per row (class or sample)."""
        R = ehs.shape[0]
        ctxp = E.linear(ctx, ehs, pk.ehp_w, pk.ehp_w.shape[0], bias=pk.ehp_b)
        v_all = E.linear(ctx, ctxp, pk.xv_w, pk.xv_total)
        table = torch.empty(R, pk.xv_total, device=ctx.device, dtype=torch.float32)
        for t in self._transformers():
            q = pk.tr[id(t)]
            E.linear(ctx, v_all, q.xo_w, q.ch, K=q.ch, c_off=q.xv_off, bias=q.xo_b,
                     out=table[:, q.xv_off:], out_ld=pk.xv_total)
        return table

    def run(self, ctx, pk, a_in, t, U, rep, H, W, xattn, xattn_idx=None, mse=None, share_prefix=False):
        """a_in: staged conv_in operand [S*H*W, kpad] ([U*H*W, kpad] with ``share_prefix``); t: [U] fp32 noise labels
        (sample s = u*rep + r).  Returns the NHWC prediction [S*H*W, Cout] (fp32) or, with ``mse``, fills mse['err']
        and returns None."""
        S = U * rep
        boc = self.config.block_out_channels
        div = rep if (share_prefix and rep > 1) else 1
        assert a_in.shape[0] == (U if div > 1 else S) * H * W
        # time embedding: sincos -> MLP -> SiLU (every consumer applies SiLU first) -> all resnets' projections
        te = E.timestep_embed(ctx, t, U, rep, boc[0], self.config.freq_shift)
        e1 = E.linear(ctx, te, pk.te1_w, pk.te1_w.shape[0], bias=pk.te1_b, act=L.ACT_SILU)
        e2 = E.linear(ctx, e1, pk.te2_w, pk.te2_w.shape[0], bias=pk.te2_b, act=L.ACT_SILU)
        temb = E.linear(ctx, e2, pk.temb_w, pk.temb_total, bias=pk.temb_b, out_dtype=torch.float32)

        h = _Act(E.linear(ctx, a_in, pk.conv_in_w, boc[0], bias=pk.conv_in_b, k_alg=9 * self.config.in_channels,
                          gn_stats=True), boc[0], div)
        skips = [h]
        for i, blk in enumerate(self.down_blocks):
            for j, r in enumerate(blk.resnets):
                h = self._resnet(ctx, pk.res[id(r)], temb, h, None, U, rep, H, W)
                if hasattr(blk, "attentions"):
                    h = self._transformer(ctx, pk.tr[id(blk.attentions[j])], xattn, xattn_idx, h, U, rep, H, W)
                skips.append(h)
            if hasattr(blk, "downsamplers"):
                sp = pk.samp[id(blk.downsamplers[0])]
                NB = U if h.div > 1 else S
                h = _Act(E.gemm(ctx, E.conv3x3_segs(h.t, h.C, H, W, stride=2), sp.w, h.C, NB, H // 2, W // 2, bias=sp.b,
                                gn_stats=True), h.C, h.div)
                H, W = H // 2, W // 2
                skips.append(h)
        mb = self.mid_block
        h = self._resnet(ctx, pk.res[id(mb.resnets[0])], temb, h, None, U, rep, H, W)
        h = self._transformer(ctx, pk.tr[id(mb.attentions[0])], xattn, xattn_idx, h, U, rep, H, W)
        h = self._resnet(ctx, pk.res[id(mb.resnets[1])], temb, h, None, U, rep, H, W)
        for i, blk in enumerate(self.up_blocks):
            for j, r in enumerate(blk.resnets):
                h = self._resnet(ctx, pk.res[id(r)], temb, h, skips.pop(), U, rep, H, W)  # cat folded into GN + K segs
                if hasattr(blk, "attentions"):
                    h = self._transformer(ctx, pk.tr[id(blk.attentions[j])], xattn, xattn_idx, h, U, rep, H, W)
            if hasattr(blk, "upsamplers"):
                sp = pk.samp[id(blk.upsamplers[0])]
                if E.FOLD_UPSAMPLE:
                    h = _Act(E.upsample_conv(ctx, h.t, sp.wph, sp.b, h.C, h.C, S, H, W), h.C, 1)
                    H, W = 2 * H, 2 * W
                else:
                    up = E.upsample2x(ctx, h.t, S, H, W, h.C)
                    H, W = 2 * H, 2 * W
                    h = _Act(E.gemm(ctx, E.conv3x3_segs(up, h.C, H, W), sp.w, h.C, S, H, W, bias=sp.b, gn_stats=True),
                             h.C, 1)
        assert h.div == 1
        Ch = h.C
        Co = self.config.out_channels
        gn = (ctx, h.t, Ch, None, 0, S, H, W, pk.out_g, pk.out_bn, self.config.norm_eps, True, pk.out_w, Co)
        if mse is not None and mse.get("fused", False):
            E.gn_conv3x3(*gn, st0=h.st, bias=pk.out_b, mse=mse, want_out=False)
            return None
        pred = E.gn_conv3x3(*gn, st0=h.st, bias=pk.out_b, out_dtype=torch.float32)
        if mse is not None:
            E.eps_mse(ctx, pred, mse["target"], mse.get("scale"), S, mse.get("div", 1), H * W * Co, mse["err"])
            return None
        return pred

    def make_ctx(self, device):
        return E.Ctx(device=device, precision=self.precision)

    # ---- reference-compatible forward (nets/unet.py:186-195) -------------------------------------------------
    @torch.no_grad()
    def forward(self, x, noise_labels, downblock_additional_residuals=None, midblock_additional_residuals=None,
                encoder_hidden_states=None):
        if downblock_additional_residuals is not None or midblock_additional_residuals is not None:
            raise NotImplementedError("ControlNet-style additional residuals are not used by the reference pipeline")
        if not x.is_cuda:
            raise RuntimeError("dcb200.UNetCondition2D runs on a CUDA (sm_100a) device only; there is no CPU path")
        ehs = encoder_hidden_states
        if ehs is None or ehs.dim() != 3 or ehs.shape[1] != 1:
            raise NotImplementedError("encoder_hidden_states must be [B,1,hid] (the reference's class-token form)")
        B, Cin, H, W = x.shape
        ctx = self.make_ctx(x.device)
        pk = self.packed(ctx)
        t = noise_labels.to(x.device, torch.float32).reshape(-1).expand(B).contiguous()
        a_in, _ = E.prologue(ctx, 0, x.contiguous().float(), B, 1, Cin, H, W, pk.kpad_in)
        xattn = self.cross_attn_table(ctx, pk, E.cast(ctx, ehs.reshape(B, -1)))
        pred = self.run(ctx, pk, a_in, t, B, 1, H, W, xattn)
        Co = self.config.out_channels
        return E.nhwc_to_nchw(ctx, pred, B, H * W, Co, Co).reshape(B, Co, H, W)
