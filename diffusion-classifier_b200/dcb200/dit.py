"""DiT -- drop-in for the reference's nets/dit.py:8-51 (kwargs passthrough to diffusers.DiTTransformer2DModel
0.31.0), executed by libdcb200's sm_100a kernels.  Same constructor, same ``forward(x, noise_labels,
encoder_hidden_states=None)`` (the third argument carries the int64 class labels, nets/dit.py:49-51) and the same
state_dict keys as diffusers.  Eval semantics only (no label dropout), see DESIGN.md.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Optional

import numpy as np
import torch
import torch.nn as nn

from . import _lib as L
from . import engine as E
from .unet import _P, _Attn, _TimeEmb


LOG2E = 1.4426950408889634
FOLD_NORMS = __import__("os").environ.get("DCB_FOLD_NORMS", "1") != "0"   # A/B switch: attention norm pre-pass in the QKV epilogue


def _sincos_1d(dim, pos):
    omega = np.arange(dim // 2, dtype=np.float64) / (dim / 2.0)
    omega = 1.0 / 10000 ** omega
    out = np.einsum("m,d->md", pos.reshape(-1), omega)
    return np.concatenate([np.sin(out), np.cos(out)], axis=1)


def sincos_2d(dim, g):
    """diffusers get_2d_sincos_pos_embed(dim, g, base_size=g): note meshgrid(w, h) -- the 'h' half uses x coords."""
    gh = np.arange(g, dtype=np.float32)
    gw = np.arange(g, dtype=np.float32)
    grid = np.stack(np.meshgrid(gw, gh), axis=0).reshape([2, 1, g, g])
    return np.concatenate([_sincos_1d(dim // 2, grid[0]), _sincos_1d(dim // 2, grid[1])], axis=1)


class _PatchEmbed(_P):
    def __init__(self, size, p, cin, D):
        super().__init__()
        self.proj = nn.Conv2d(cin, D, kernel_size=p, stride=p, bias=True)
        self.register_buffer("pos_embed", torch.from_numpy(sincos_2d(D, size // p)).float().unsqueeze(0),
                             persistent=False)


class _LabelEmb(_P):
    def __init__(self, n, D):
        super().__init__()
        self.embedding_table = nn.Embedding(n + 1, D)  # +1: cfg/null row (class_dropout_prob = 0.1 > 0)


class _CombinedEmb(_P):
    def __init__(self, n, D):
        super().__init__()
        self.timestep_embedder = _TimeEmb(256, D)
        self.class_embedder = _LabelEmb(n, D)


class _AdaLNZero(_P):
    def __init__(self, D, n):
        super().__init__()
        self.emb = _CombinedEmb(n, D)
        self.linear = nn.Linear(D, 6 * D)


class _GELU(_P):
    def __init__(self, d, inner):
        super().__init__()
        self.proj = nn.Linear(d, inner)


class _FFGelu(_P):
    def __init__(self, d):
        super().__init__()
        self.net = nn.ModuleList([_GELU(d, 4 * d), nn.Dropout(0.0), nn.Linear(4 * d, d)])


class _DiTBlock(_P):
    def __init__(self, D, n_emb, bias):
        super().__init__()
        self.norm1 = _AdaLNZero(D, n_emb)
        self.attn1 = _Attn(D, D, bias=bias)
        self.ff = _FFGelu(D)


class DiT(nn.Module):
    def __init__(
        self,
        num_attention_heads: int = 16,
        attention_head_dim: int = 72,
        in_channels: int = 4,
        out_channels: Optional[int] = None,
        num_layers: int = 28,
        dropout: float = 0.0,
        norm_num_groups: int = 32,
        attention_bias: bool = True,
        sample_size: int = 32,
        patch_size: int = 2,
        activation_fn: str = "gelu-approximate",
        num_embeds_ada_norm: Optional[int] = 1000,
        upcast_attention: bool = False,
        norm_type: str = "ada_norm_zero",
        norm_elementwise_affine: bool = False,
        norm_eps: float = 1e-5,
    ):
        super().__init__()
        if norm_type != "ada_norm_zero" or activation_fn != "gelu-approximate" or norm_elementwise_affine \
                or dropout != 0.0 or upcast_attention or not attention_bias:
            raise NotImplementedError("dcb200.DiT implements the DiTTransformer2DModel configuration the reference uses")
        D = num_attention_heads * attention_head_dim
        out_channels = in_channels if out_channels is None else out_channels
        if attention_head_dim not in (32, 64, 96, 128) or D % 64 or D > 1024:
            raise NotImplementedError(f"head_dim {attention_head_dim} / width {D} outside the kernel set")
        self.config = SimpleNamespace(
            num_attention_heads=num_attention_heads, attention_head_dim=attention_head_dim, in_channels=in_channels,
            out_channels=out_channels, num_layers=num_layers, sample_size=sample_size, patch_size=patch_size,
            num_embeds_ada_norm=num_embeds_ada_norm, norm_eps=norm_eps)
        self.pos_embed = _PatchEmbed(sample_size, patch_size, in_channels, D)
        self.transformer_blocks = nn.ModuleList(
            [_DiTBlock(D, num_embeds_ada_norm, attention_bias) for _ in range(num_layers)])
        self.proj_out_1 = nn.Linear(D, 2 * D)
        self.proj_out_2 = nn.Linear(D, patch_size * patch_size * out_channels)
        self.D = D
        self._packs = {}
        self.precision = "bf16"

    def _version(self):
        return E.params_version(self)

    @staticmethod
    def _fold_qscale(ctx):
        # bf16 product path only: the fp32-verify engine keeps the reference's order of operations (1e-4 mode)
        return ctx.precision == "bf16"

    def packed(self, ctx):
        key = (ctx.precision, str(ctx.device))
        ver = self._version()
        hit = self._packs.get(key)
        if hit is not None and hit.version == ver:
            return hit
        pk = self._pack(ctx)
        pk.version = ver
        pk.gen = next(E.PACK_GEN)
        self._packs[key] = pk
        return pk

    @torch.no_grad()
    def _pack(self, ctx):
        def w(t):
            return E.cast(ctx, t)

        def f32(t):
            return t.detach().to(ctx.device, torch.float32).contiguous()

        cfg, D = self.config, self.D
        p, cin = cfg.patch_size, cfg.in_channels
        pk = SimpleNamespace()
        pk.kpad_in = (p * p * cin + 63) // 64 * 64
        pk.pe_w = E.pack_conv(ctx, self.pos_embed.proj.weight, pk.kpad_in)     # K order (py, px, c), zero padded
        pk.pe_b = f32(self.pos_embed.proj.bias)
        pk.pos = w(self.pos_embed.pos_embed[0])
        blks = list(self.transformer_blocks)
        pk.te1_w = E.pack_rows(ctx, [b.norm1.emb.timestep_embedder.linear_1.weight for b in blks], 0)
        pk.te1_b = f32(torch.cat([b.norm1.emb.timestep_embedder.linear_1.bias.detach() for b in blks], 0))
        pk.blk = []
        for b in blks:
            q = SimpleNamespace()
            te = b.norm1.emb.timestep_embedder
            q.te2_w, q.te2_b = w(te.linear_2.weight), f32(te.linear_2.bias)
            q.table = w(b.norm1.emb.class_embedder.embedding_table.weight)
            q.mod_w, q.mod_b = w(b.norm1.linear.weight), f32(b.norm1.linear.bias)
            a = b.attn1
            # the softmax exponent scale c = d^-0.5 log2(e) is folded into the query projection (weights and bias, in fp32,
            # before the one rounding to the engine dtype): QK^T then IS the base-2 exponent and the single-pass attention
            # kernel spends one instruction less per score (attention_tc.cu); attention() is told scale = 1 / log2(e)
            cq = (cfg.attention_head_dim ** -0.5) * LOG2E if self._fold_qscale(ctx) else 1.0
            q.qkv_w = E.pack_rows(ctx, [a.to_q.weight.detach() * cq, a.to_k.weight, a.to_v.weight], 0)
            q.qkv_b = f32(torch.cat([a.to_q.bias.detach() * cq, a.to_k.bias.detach(), a.to_v.bias.detach()], 0))
            q.o_w, q.o_b = w(a.to_out[0].weight), f32(a.to_out[0].bias)
            q.f1_w, q.f1_b = w(b.ff.net[0].proj.weight), f32(b.ff.net[0].proj.bias)
            q.f2_w, q.f2_b = w(b.ff.net[2].weight), f32(b.ff.net[2].bias)
            pk.blk.append(q)
        pk.po1_w, pk.po1_b = w(self.proj_out_1.weight), f32(self.proj_out_1.bias)
        pk.po2_w, pk.po2_b = w(self.proj_out_2.weight), f32(self.proj_out_2.bias)
        return pk

    def run(self, ctx, pk, a_in, t, U, rep, labels, mse=None):
        """a_in: patchified operand [S*N, kpad]; t: [U] fp32; labels: [S] int32 class ids (sample s = u*rep + r)."""
        cfg, D = self.config, self.D
        S = U * rep
        g = cfg.sample_size // cfg.patch_size
        N = g * g
        heads, hd = cfg.num_attention_heads, cfg.attention_head_dim
        h = E.linear(ctx, a_in, pk.pe_w, D, bias=pk.pe_b, residual=pk.pos, res_ld=D, res_mod=N)
        tproj = E.timestep_embed(ctx, t, U, rep, 256, 1.0)
        c1_all = E.linear(ctx, tproj, pk.te1_w, pk.te1_w.shape[0], bias=pk.te1_b, act=L.ACT_SILU)
        sc0 = None
        for i, q in enumerate(pk.blk):
            # SiLU(timestep_emb + class_emb): every consumer of the conditioning applies SiLU first
            sc = E.linear(ctx, c1_all, q.te2_w, D, K=D, c_off=i * D, bias=q.te2_b, residual=q.table, res_ld=D,
                          res_idx=labels, act_post=L.ACT_SILU)
            if i == 0:
                sc0 = sc
            mod = E.linear(ctx, sc, q.mod_w, 6 * D, bias=q.mod_b, out_dtype=torch.float32)
            n = E.layernorm(ctx, h, None, None, 1e-6, scale=mod[:, D:], shift=mod, mod_ld=6 * D, rows_per_group=N)
            # the projection's epilogue leaves the per-head max |q|^2, max |k|^2 the single-pass attention kernel needs
            nws = E.attn_norms_ws(ctx, S, heads) if (FOLD_NORMS and ctx.precision == "bf16" and hd == 64 and N % 128 == 0
                                                     and N >= 1024) else None
            qkv = E.linear(ctx, n, q.qkv_w, 3 * D, bias=q.qkv_b, attn_norms=None if nws is None else (nws, heads, N))
            att = E.attention(ctx, qkv, S, N, heads, hd, scale=(1.0 / LOG2E) if self._fold_qscale(ctx) else None,
                              norms_ready=nws)
            h = E.linear(ctx, att, q.o_w, D, bias=q.o_b, gate=mod[:, 2 * D:], gate_ld=6 * D, rows_per_group=N,
                         residual=h, res_ld=D)
            n = E.layernorm(ctx, h, None, None, cfg.norm_eps, scale=mod[:, 4 * D:], shift=mod[:, 3 * D:], mod_ld=6 * D,
                            rows_per_group=N)
            f = E.linear(ctx, n, q.f1_w, 4 * D, bias=q.f1_b, act=L.ACT_GELU_TANH)
            h = E.linear(ctx, f, q.f2_w, D, bias=q.f2_b, gate=mod[:, 5 * D:], gate_ld=6 * D, rows_per_group=N,
                         residual=h, res_ld=D)
        so = E.linear(ctx, sc0, pk.po1_w, 2 * D, bias=pk.po1_b, out_dtype=torch.float32)
        n = E.layernorm(ctx, h, None, None, 1e-6, scale=so[:, D:], shift=so, mod_ld=2 * D, rows_per_group=N)
        No = cfg.patch_size ** 2 * cfg.out_channels
        if mse is not None and mse.get("fused", False):
            E.gemm(ctx, [E.seg(n, D, 1, N)], pk.po2_w, No, S, 1, N, bias=pk.po2_b, mse=mse, want_out=False)
            return None
        pred = E.gemm(ctx, [E.seg(n, D, 1, N)], pk.po2_w, No, S, 1, N, bias=pk.po2_b, out_dtype=torch.float32)
        if mse is not None:
            E.eps_mse(ctx, pred, mse["target"], mse.get("scale"), S, mse.get("div", 1), N * No, mse["err"])
            return None
        return pred

    def make_ctx(self, device):
        return E.Ctx(device=device, precision=self.precision)

    @torch.no_grad()
    def forward(self, x, noise_labels, encoder_hidden_states=None):
        if not x.is_cuda:
            raise RuntimeError("dcb200.DiT runs on a CUDA (sm_100a) device only; there is no CPU path")
        if encoder_hidden_states is None:
            raise ValueError("class labels are required (ada_norm_zero conditioning)")
        cfg = self.config
        B, Cin, H, W = x.shape
        ctx = self.make_ctx(x.device)
        pk = self.packed(ctx)
        t = noise_labels.to(x.device, torch.float32).reshape(-1).expand(B).contiguous()
        labels = encoder_hidden_states.to(x.device, torch.int32).reshape(-1).contiguous()
        a_in, _ = E.prologue(ctx, 1, x.contiguous().float(), B, 1, Cin, H, W, pk.kpad_in, patch=cfg.patch_size)
        pred = self.run(ctx, pk, a_in, t, B, 1, labels)
        g, p = cfg.sample_size // cfg.patch_size, cfg.patch_size
        return E.unpatchify(ctx, pred, B, g, p, cfg.out_channels, p * p * cfg.out_channels)
