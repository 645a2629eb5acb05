"""ORACLE (test infrastructure only -- never imported by the product path).

fp32 PyTorch restatement of the third-party arithmetic the reference delegates to:
``diffusers==0.31.0`` (pinned in /root/reference/requirements.txt:9), which is NOT vendored
under /root/reference and NOT installable here (no network).  The reference's own call sites
that select these code paths are

  * nets/unet.py:2,77,134-183,186-195  -> diffusers.UNet2DConditionModel
  * nets/dit.py:2,8,29-46,49-51        -> diffusers.DiTTransformer2DModel

Only the branches that the reference's kwargs select are restated.
Parameter names follow the diffusers state_dict schema so that a
checkpoint trained with the reference loads into these modules and into the product modules.

PARITY UNPINNED: the reference ships no tests / golden vectors for the denoisers and diffusers
itself cannot be run here; this restatement (plus the closed-form KATs in tests/) *is* the pin.
The loop around the denoiser (classify / schedule / q_sample) is pinned against the reference's
verbatim code, see oracle/reference_loader.py and oracle/make_golden.py.
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


# --------------------------------------------------------------------------------------------
# embeddings.py :: get_timestep_embedding / Timesteps / TimestepEmbedding
# --------------------------------------------------------------------------------------------
def get_timestep_embedding(timesteps, embedding_dim, flip_sin_to_cos=False, downscale_freq_shift=1.0,
                           scale=1.0, max_period=10000):
    half = embedding_dim // 2
    exponent = -math.log(max_period) * torch.arange(half, dtype=torch.float32, device=timesteps.device)
    exponent = exponent / (half - downscale_freq_shift)
    emb = torch.exp(exponent)
    emb = timesteps[:, None].float() * emb[None, :]
    emb = scale * emb
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
    if embedding_dim % 2 == 1:
        emb = F.pad(emb, (0, 1, 0, 0))
    return emb


class TimestepEmbedding(nn.Module):
    def __init__(self, in_channels, time_embed_dim):
        super().__init__()
        self.linear_1 = nn.Linear(in_channels, time_embed_dim)
        self.linear_2 = nn.Linear(time_embed_dim, time_embed_dim)

    def forward(self, x):
        return self.linear_2(F.silu(self.linear_1(x)))


# --------------------------------------------------------------------------------------------
# attention_processor.py :: Attention + AttnProcessor2_0 (SDPA), attention.py :: FeedForward
# --------------------------------------------------------------------------------------------
class _ToOut(nn.ModuleList):
    pass


class Attention(nn.Module):
    def __init__(self, query_dim, heads, dim_head, cross_attention_dim=None, bias=False):
        super().__init__()
        inner = heads * dim_head
        ctx = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.heads = heads
        self.to_q = nn.Linear(query_dim, inner, bias=bias)
        self.to_k = nn.Linear(ctx, inner, bias=bias)
        self.to_v = nn.Linear(ctx, inner, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim, bias=True), nn.Dropout(0.0)])

    def forward(self, x, encoder_hidden_states=None):
        ctx = x if encoder_hidden_states is None else encoder_hidden_states
        B, N, _ = x.shape
        q, k, v = self.to_q(x), self.to_k(ctx), self.to_v(ctx)
        d = q.shape[-1] // self.heads
        q = q.view(B, -1, self.heads, d).transpose(1, 2)
        k = k.view(B, -1, self.heads, d).transpose(1, 2)
        v = v.view(B, -1, self.heads, d).transpose(1, 2)
        # explicit softmax(QK^T/sqrt(d))V in fp32: mathematically what SDPA computes
        s = torch.matmul(q, k.transpose(-1, -2)) * (1.0 / math.sqrt(d))
        o = torch.matmul(torch.softmax(s, dim=-1), v)
        o = o.transpose(1, 2).reshape(B, N, self.heads * d)
        return self.to_out[0](o)


class GEGLU(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        h, gate = self.proj(x).chunk(2, dim=-1)
        return h * F.gelu(gate)


class GELUProj(nn.Module):
    def __init__(self, dim_in, dim_out, approximate="tanh"):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out)
        self.approximate = approximate

    def forward(self, x):
        return F.gelu(self.proj(x), approximate=self.approximate)


class FeedForward(nn.Module):
    def __init__(self, dim, activation_fn):
        super().__init__()
        inner = dim * 4
        act = GEGLU(dim, inner) if activation_fn == "geglu" else GELUProj(dim, inner, "tanh")
        self.net = nn.ModuleList([act, nn.Dropout(0.0), nn.Linear(inner, dim)])

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


# --------------------------------------------------------------------------------------------
# attention.py :: BasicTransformerBlock (norm_type='layer_norm' branch) + transformer_2d.py
# --------------------------------------------------------------------------------------------
class BasicTransformerBlockLN(nn.Module):
    def __init__(self, dim, heads, dim_head, cross_attention_dim):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-5)
        self.attn1 = Attention(dim, heads, dim_head, None, bias=False)
        self.norm2 = nn.LayerNorm(dim, eps=1e-5)
        self.attn2 = Attention(dim, heads, dim_head, cross_attention_dim, bias=False)
        self.norm3 = nn.LayerNorm(dim, eps=1e-5)
        self.ff = FeedForward(dim, "geglu")

    def forward(self, h, ehs):
        h = self.attn1(self.norm1(h)) + h
        h = self.attn2(self.norm2(h), ehs) + h
        h = self.ff(self.norm3(h)) + h
        return h


class Transformer2DModel(nn.Module):
    def __init__(self, heads, dim_head, in_channels, cross_attention_dim, norm_num_groups=32):
        super().__init__()
        inner = heads * dim_head
        self.norm = nn.GroupNorm(norm_num_groups, in_channels, eps=1e-6, affine=True)
        self.proj_in = nn.Conv2d(in_channels, inner, 1)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlockLN(inner, heads, dim_head, cross_attention_dim)])
        self.proj_out = nn.Conv2d(inner, in_channels, 1)

    def forward(self, x, ehs):
        B, C, H, W = x.shape
        r = x
        h = self.proj_in(self.norm(x))
        h = h.permute(0, 2, 3, 1).reshape(B, H * W, -1)
        for blk in self.transformer_blocks:
            h = blk(h, ehs)
        h = h.reshape(B, H, W, -1).permute(0, 3, 1, 2)
        return self.proj_out(h) + r


# --------------------------------------------------------------------------------------------
# resnet.py :: ResnetBlock2D, downsampling.py / upsampling.py
# --------------------------------------------------------------------------------------------
class ResnetBlock2D(nn.Module):
    def __init__(self, in_channels, out_channels, temb_channels, groups=32, eps=1e-5):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, in_channels, eps=eps, affine=True)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, 1, 1)
        self.time_emb_proj = nn.Linear(temb_channels, out_channels)
        self.norm2 = nn.GroupNorm(groups, out_channels, eps=eps, affine=True)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, 1, 1)
        self.conv_shortcut = nn.Conv2d(in_channels, out_channels, 1) if in_channels != out_channels else None

    def forward(self, x, temb):
        h = self.conv1(F.silu(self.norm1(x)))
        h = h + self.time_emb_proj(F.silu(temb))[:, :, None, None]
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class Downsample2D(nn.Module):
    def __init__(self, ch):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 3, stride=2, padding=1)

    def forward(self, x):
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, ch):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class _Block(nn.Module):
    """Down / up / mid block container with the diffusers attribute names."""

    def __init__(self):
        super().__init__()


def _as_tuple(v, n):
    return tuple(v) if isinstance(v, (tuple, list)) else (v,) * n


# --------------------------------------------------------------------------------------------
# unets/unet_2d_condition.py :: UNet2DConditionModel (subset selected by nets/unet.py:134-183)
# --------------------------------------------------------------------------------------------
class UNet2DConditionModel(nn.Module):
    def __init__(self, sample_size=None, in_channels=4, out_channels=4, center_input_sample=False,
                 flip_sin_to_cos=True, freq_shift=0,
                 down_block_types=("CrossAttnDownBlock2D",) * 3 + ("DownBlock2D",),
                 mid_block_type="UNetMidBlock2DCrossAttn",
                 up_block_types=("UpBlock2D",) + ("CrossAttnUpBlock2D",) * 3,
                 only_cross_attention=False, block_out_channels=(320, 640, 1280, 1280), layers_per_block=2,
                 downsample_padding=1, mid_block_scale_factor=1, dropout=0.0, act_fn="silu",
                 norm_num_groups=32, norm_eps=1e-5, cross_attention_dim=1280,
                 transformer_layers_per_block=1, reverse_transformer_layers_per_block=None,
                 encoder_hid_dim=None, encoder_hid_dim_type=None, attention_head_dim=8,
                 num_attention_heads=None, **unused):
        super().__init__()
        assert mid_block_type == "UNetMidBlock2DCrossAttn" and act_fn == "silu"
        assert transformer_layers_per_block == 1 and not center_input_sample and dropout == 0.0
        assert encoder_hid_dim_type == "text_proj" and downsample_padding == 1
        for k, v in unused.items():  # every other kwarg must be at the wrapper default
            assert v in (None, False, "default", "positional", 1.0, 3, 64), (k, v)
        boc = tuple(block_out_channels)
        n = len(boc)
        lpb = _as_tuple(layers_per_block, n)
        heads = num_attention_heads or attention_head_dim  # diffusers naming quirk -> 8 heads
        heads = _as_tuple(heads, n)
        self.config = SimpleNamespace(
            sample_size=sample_size, in_channels=in_channels, out_channels=out_channels,
            block_out_channels=boc, layers_per_block=layers_per_block, encoder_hid_dim=encoder_hid_dim,
            cross_attention_dim=cross_attention_dim, down_block_types=tuple(down_block_types),
            up_block_types=tuple(up_block_types), flip_sin_to_cos=flip_sin_to_cos, freq_shift=freq_shift)
        tdim = boc[0] * 4
        self.conv_in = nn.Conv2d(in_channels, boc[0], 3, padding=1)
        self.time_embedding = TimestepEmbedding(boc[0], tdim)
        self.encoder_hid_proj = nn.Linear(encoder_hid_dim, cross_attention_dim)

        self.down_blocks = nn.ModuleList()
        out = boc[0]
        for i, typ in enumerate(down_block_types):
            inp, out = out, boc[i]
            blk = _Block()
            blk.resnets = nn.ModuleList(
                [ResnetBlock2D(inp if j == 0 else out, out, tdim, norm_num_groups, norm_eps) for j in range(lpb[i])])
            if typ == "CrossAttnDownBlock2D":
                blk.attentions = nn.ModuleList(
                    [Transformer2DModel(heads[i], out // heads[i], out, cross_attention_dim, norm_num_groups)
                     for _ in range(lpb[i])])
            else:
                assert typ == "DownBlock2D", typ
            if i != n - 1:
                blk.downsamplers = nn.ModuleList([Downsample2D(out)])
            self.down_blocks.append(blk)

        C = boc[-1]
        self.mid_block = _Block()
        self.mid_block.attentions = nn.ModuleList(
            [Transformer2DModel(heads[-1], C // heads[-1], C, cross_attention_dim, norm_num_groups)])
        self.mid_block.resnets = nn.ModuleList(
            [ResnetBlock2D(C, C, tdim, norm_num_groups, norm_eps) for _ in range(2)])

        self.up_blocks = nn.ModuleList()
        rb, rl, rh = boc[::-1], lpb[::-1], heads[::-1]
        out = rb[0]
        for i, typ in enumerate(up_block_types):
            prev, out = out, rb[i]
            inn = rb[min(i + 1, n - 1)]
            L = rl[i] + 1
            blk = _Block()
            res = []
            for j in range(L):
                skip_c = inn if j == L - 1 else out
                rin = prev if j == 0 else out
                res.append(ResnetBlock2D(rin + skip_c, out, tdim, norm_num_groups, norm_eps))
            blk.resnets = nn.ModuleList(res)
            if typ == "CrossAttnUpBlock2D":
                blk.attentions = nn.ModuleList(
                    [Transformer2DModel(rh[i], out // rh[i], out, cross_attention_dim, norm_num_groups)
                     for _ in range(L)])
            else:
                assert typ == "UpBlock2D", typ
            if i != n - 1:
                blk.upsamplers = nn.ModuleList([Upsample2D(out)])
            self.up_blocks.append(blk)

        self.conv_norm_out = nn.GroupNorm(norm_num_groups, boc[0], eps=norm_eps)
        self.conv_out = nn.Conv2d(boc[0], out_channels, 3, padding=1)

    def forward(self, sample, timestep, encoder_hidden_states=None, down_block_additional_residuals=None,
                mid_block_additional_residual=None, return_dict=True, **kw):
        assert down_block_additional_residuals is None and mid_block_additional_residual is None
        cfg = self.config
        if not torch.is_tensor(timestep):
            timestep = torch.tensor([timestep], device=sample.device)
        timestep = timestep.reshape(-1).expand(sample.shape[0])
        t_emb = get_timestep_embedding(timestep, cfg.block_out_channels[0], cfg.flip_sin_to_cos, cfg.freq_shift)
        emb = self.time_embedding(t_emb.to(sample.dtype))
        ehs = self.encoder_hid_proj(encoder_hidden_states)
        h = self.conv_in(sample)
        skips = [h]
        for blk in self.down_blocks:
            for j, res in enumerate(blk.resnets):
                h = res(h, emb)
                if hasattr(blk, "attentions"):
                    h = blk.attentions[j](h, ehs)
                skips.append(h)
            if hasattr(blk, "downsamplers"):
                h = blk.downsamplers[0](h)
                skips.append(h)
        h = self.mid_block.resnets[0](h, emb)
        h = self.mid_block.attentions[0](h, ehs)
        h = self.mid_block.resnets[1](h, emb)
        for blk in self.up_blocks:
            for j, res in enumerate(blk.resnets):
                h = torch.cat([h, skips.pop()], dim=1)
                h = res(h, emb)
                if hasattr(blk, "attentions"):
                    h = blk.attentions[j](h, ehs)
            if hasattr(blk, "upsamplers"):
                h = blk.upsamplers[0](h)
        h = self.conv_out(F.silu(self.conv_norm_out(h)))
        return (h,)


class UNet2DModel(nn.Module):
    """nets/unet.py:10-71 also subclasses diffusers.UNet2DModel (unused by any experiment)."""

    def __init__(self, *a, **k):
        super().__init__()
        raise NotImplementedError("UNet2D (unconditional) is not on the classify hot path")


# --------------------------------------------------------------------------------------------
# embeddings.py :: 2-D sincos positional embedding, PatchEmbed, CombinedTimestepLabelEmbeddings
# --------------------------------------------------------------------------------------------
def _sincos_1d(embed_dim, pos):
    omega = np.arange(embed_dim // 2, dtype=np.float64) / (embed_dim / 2.0)
    omega = 1.0 / 10000 ** omega
    out = np.einsum("m,d->md", pos.reshape(-1), omega)
    return np.concatenate([np.sin(out), np.cos(out)], axis=1)


def get_2d_sincos_pos_embed(embed_dim, grid_size, base_size, interpolation_scale=1.0):
    gh = np.arange(grid_size, dtype=np.float32) / (grid_size / base_size) / interpolation_scale
    gw = np.arange(grid_size, dtype=np.float32) / (grid_size / base_size) / interpolation_scale
    grid = np.stack(np.meshgrid(gw, gh), axis=0).reshape([2, 1, grid_size, grid_size])  # w goes first
    emb_h = _sincos_1d(embed_dim // 2, grid[0])
    emb_w = _sincos_1d(embed_dim // 2, grid[1])
    return np.concatenate([emb_h, emb_w], axis=1)


class PatchEmbed(nn.Module):
    def __init__(self, size, patch_size, in_channels, embed_dim):
        super().__init__()
        self.proj = nn.Conv2d(in_channels, embed_dim, kernel_size=patch_size, stride=patch_size, bias=True)
        g = size // patch_size
        pe = get_2d_sincos_pos_embed(embed_dim, g, base_size=g)
        self.register_buffer("pos_embed", torch.from_numpy(pe).float().unsqueeze(0), persistent=False)

    def forward(self, x):
        x = self.proj(x).flatten(2).transpose(1, 2)
        return (x + self.pos_embed).to(x.dtype)


class LabelEmbedding(nn.Module):
    def __init__(self, num_classes, hidden, dropout_prob=0.1):
        super().__init__()
        self.embedding_table = nn.Embedding(num_classes + int(dropout_prob > 0), hidden)

    def forward(self, labels):
        # eval semantics only: label dropout is a training-time op ("EMA train/eval quirk")
        return self.embedding_table(labels)


class CombinedTimestepLabelEmbeddings(nn.Module):
    def __init__(self, num_classes, dim):
        super().__init__()
        self.timestep_embedder = TimestepEmbedding(256, dim)
        self.class_embedder = LabelEmbedding(num_classes, dim)

    def forward(self, timestep, labels, hidden_dtype=torch.float32):
        tp = get_timestep_embedding(timestep, 256, flip_sin_to_cos=True, downscale_freq_shift=1)
        return self.timestep_embedder(tp.to(hidden_dtype)) + self.class_embedder(labels)


class AdaLayerNormZero(nn.Module):
    def __init__(self, dim, num_embeddings):
        super().__init__()
        self.emb = CombinedTimestepLabelEmbeddings(num_embeddings, dim)
        self.linear = nn.Linear(dim, 6 * dim)
        self.norm = nn.LayerNorm(dim, elementwise_affine=False, eps=1e-6)

    def forward(self, x, timestep, labels):
        e = self.linear(F.silu(self.emb(timestep, labels, x.dtype)))
        sh_a, sc_a, g_a, sh_m, sc_m, g_m = e.chunk(6, dim=1)
        return self.norm(x) * (1 + sc_a[:, None]) + sh_a[:, None], g_a, sh_m, sc_m, g_m


class BasicTransformerBlockAdaZero(nn.Module):
    def __init__(self, dim, heads, dim_head, num_embeds, norm_eps, attention_bias):
        super().__init__()
        self.norm1 = AdaLayerNormZero(dim, num_embeds)
        self.attn1 = Attention(dim, heads, dim_head, None, bias=attention_bias)
        self.norm3 = nn.LayerNorm(dim, eps=norm_eps, elementwise_affine=False)
        self.ff = FeedForward(dim, "gelu-approximate")

    def forward(self, h, timestep, labels):
        n, g_a, sh_m, sc_m, g_m = self.norm1(h, timestep, labels)
        h = g_a.unsqueeze(1) * self.attn1(n) + h
        n = self.norm3(h) * (1 + sc_m[:, None]) + sh_m[:, None]
        h = g_m.unsqueeze(1) * self.ff(n) + h
        return h


# --------------------------------------------------------------------------------------------
# transformers/dit_transformer_2d.py :: DiTTransformer2DModel
# --------------------------------------------------------------------------------------------
class DiTTransformer2DModel(nn.Module):
    def __init__(self, num_attention_heads=16, attention_head_dim=72, in_channels=4, out_channels=None,
                 num_layers=28, dropout=0.0, norm_num_groups=32, attention_bias=True, sample_size=32,
                 patch_size=2, activation_fn="gelu-approximate", num_embeds_ada_norm=1000,
                 upcast_attention=False, norm_type="ada_norm_zero", norm_elementwise_affine=False,
                 norm_eps=1e-5):
        super().__init__()
        assert norm_type == "ada_norm_zero" and activation_fn == "gelu-approximate"
        assert not norm_elementwise_affine and dropout == 0.0
        D = num_attention_heads * attention_head_dim
        out_channels = in_channels if out_channels is None else out_channels
        self.config = SimpleNamespace(
            num_attention_heads=num_attention_heads, attention_head_dim=attention_head_dim, in_channels=in_channels,
            out_channels=out_channels, num_layers=num_layers, sample_size=sample_size, patch_size=patch_size,
            num_embeds_ada_norm=num_embeds_ada_norm, norm_eps=norm_eps, attention_bias=attention_bias)
        self.patch_size, self.out_channels = patch_size, out_channels
        self.pos_embed = PatchEmbed(sample_size, patch_size, in_channels, D)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlockAdaZero(D, num_attention_heads, attention_head_dim, num_embeds_ada_norm,
                                          norm_eps, attention_bias) for _ in range(num_layers)])
        self.norm_out = nn.LayerNorm(D, elementwise_affine=False, eps=1e-6)
        self.proj_out_1 = nn.Linear(D, 2 * D)
        self.proj_out_2 = nn.Linear(D, patch_size * patch_size * out_channels)

    def forward(self, hidden_states, timestep=None, class_labels=None, cross_attention_kwargs=None,
                return_dict=True):
        p = self.patch_size
        g = hidden_states.shape[-1] // p
        timestep = timestep.reshape(-1).expand(hidden_states.shape[0])
        h = self.pos_embed(hidden_states)
        for blk in self.transformer_blocks:
            h = blk(h, timestep, class_labels)
        c0 = self.transformer_blocks[0].norm1.emb(timestep, class_labels, h.dtype)
        shift, scale = self.proj_out_1(F.silu(c0)).chunk(2, dim=1)
        h = self.norm_out(h) * (1 + scale[:, None]) + shift[:, None]
        h = self.proj_out_2(h)
        h = h.reshape(-1, g, g, p, p, self.out_channels)
        h = torch.einsum("nhwpqc->nchpwq", h)
        return (h.reshape(-1, self.out_channels, g * p, g * p),)
