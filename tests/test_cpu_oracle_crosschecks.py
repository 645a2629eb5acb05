"""Narrowing the "parity unpinned" part of the oracle (VERDICT r1 Next #6): diffusers 0.31.0 and pywt are not installable
offline, but INDEPENDENT third-party restatements of three of their pieces ship in this image.  Each is executed from
where it lies (nothing copied) and compared with oracle/diffusers_restated.py / oracle/haar.py and with the product's host
code:

  * TVM's port of diffusers ``get_timestep_embedding`` (tilelang/3rdparty/tvm/.../relax/frontend/nn/op.py) -- executed
    through a numpy shim of the handful of relax ops it uses;
  * MAE's ``get_2d_sincos_pos_embed`` in transformers (the function diffusers' PatchEmbed pos-embed was taken from) -- pins
    the "w goes first" meshgrid order and the [sin | cos] layout of DiT's positional table;
  * the Haar analysis bank written from pywt's documented filter coefficients (dec_lo = [1, 1]/sqrt2,
    dec_hi = [-1, 1]/sqrt2, coefficient k = sum_j f[j] x[2k + 1 - j]) as an explicit convolution + decimation."""
import ast
import contextlib
import math
import os
import sys

import numpy as np
import pytest
import torch


def _tvm_timestep_embedding_fn():
    try:
        import tilelang
    except Exception:  # pragma: no cover
        pytest.skip("tilelang (which vendors the TVM sources) is not importable")
    path = os.path.join(os.path.dirname(tilelang.__file__), "3rdparty", "tvm", "python", "tvm", "relax", "frontend", "nn",
                        "op.py")
    if not os.path.exists(path):
        pytest.skip("TVM sources not present")
    tree = ast.parse(open(path).read())
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "get_timestep_embedding"]
    if not fn:
        pytest.skip("TVM has no get_timestep_embedding")
    fn[0].returns = None
    for a in fn[0].args.args:
        a.annotation = None
    mod = ast.Module(body=[fn[0]], type_ignores=[])

    class T:                                   # stand-in for the frontend's Tensor wrapper
        def __init__(self, e):
            self._expr = np.asarray(e)

    class _nn:
        @staticmethod
        def pad(x, widths):
            return np.pad(x, ((widths[2], widths[3]), (widths[0], widths[1])))

    class _op:
        nn = _nn
        astype = staticmethod(lambda x, dt: np.asarray(x).astype(dt))
        arange = staticmethod(lambda start, end, dtype: np.arange(start, end, dtype=dtype))
        exp = staticmethod(np.exp)
        cos = staticmethod(np.cos)
        sin = staticmethod(np.sin)
        expand_dims = staticmethod(lambda x, axis: np.expand_dims(x, axis))
        concat = staticmethod(lambda xs, axis: np.concatenate(xs, axis=axis))

    class rx:
        const = staticmethod(lambda v, dt: np.asarray(v, dtype=dt))

    ns = dict(_op=_op, rx=rx, math=math, get_default_dtype=lambda: "float32", wrap_nested=lambda e, name: e, Tensor=T)
    exec(compile(mod, path, "exec"), ns)
    return lambda t, dim, **kw: ns["get_timestep_embedding"](T(t), dim, **kw)


@pytest.mark.parametrize("dim,flip,shift", [(128, True, 0), (256, True, 1), (64, False, 1), (33, True, 0)])
def test_timestep_embedding_matches_tvm_port_of_diffusers(dim, flip, shift):
    """a5.1 / a6: U-Net uses (flip_sin_to_cos=True, freq_shift=0), DiT's Timesteps(256, True, 1)."""
    from oracle import diffusers_restated as dr
    tvm_fn = _tvm_timestep_embedding_fn()
    t = np.array([-14.3, -1.7, 0.0, 0.31, 2.5, 15.0], dtype=np.float32)       # logSNR-valued noise labels
    ref = tvm_fn(t, dim, flip_sin_to_cos=flip, downscale_freq_shift=shift)
    mine = dr.get_timestep_embedding(torch.from_numpy(t), dim, flip_sin_to_cos=flip, downscale_freq_shift=shift).numpy()
    assert ref.shape == mine.shape == (6, dim)
    assert np.abs(ref - mine).max() < 2e-6


@contextlib.contextmanager
def _without_test_stubs():
    """Other tests of this suite leave spec-less stand-ins for packages this image lacks (accelerate, diffusers, ...) in
    sys.modules; transformers probes some of them with importlib.util.find_spec, which raises on a module without a spec.
    Import transformers' model files with those stand-ins out of the way."""
    names = [k for k, m in sys.modules.items()
             if k.split(".")[0] in ("accelerate", "diffusers", "comet_ml", "ema_pytorch", "pywt", "dataset")
             and getattr(m, "__spec__", None) is None]
    saved = {k: sys.modules.pop(k) for k in names}
    try:
        yield
    finally:
        sys.modules.update(saved)


def test_dit_pos_embed_matches_mae_sincos_table():
    try:
        with _without_test_stubs():
            from transformers.models.vit_mae.modeling_vit_mae import get_2d_sincos_pos_embed as mae
    except Exception:  # pragma: no cover
        pytest.skip("transformers' MAE model is not importable")
    from oracle import diffusers_restated as dr
    from dcb200.dit import sincos_2d
    for dim, g in ((768, 64), (128, 16), (64, 6)):
        ref = mae(dim, g, add_cls_token=False)
        a = dr.get_2d_sincos_pos_embed(dim, g, base_size=g)
        assert ref.shape == a.shape == (g * g, dim)
        assert np.abs(ref - a).max() < 1e-12                      # oracle == the MAE table (float64)
        assert np.abs(ref - sincos_2d(dim, g)).max() < 1e-12      # product host code == the MAE table


def _bank(x, f):
    """pywt's decimating convolution for an orthogonal 2-tap bank on an even-length signal: c[k] = sum_j f[j] x[2k+1-j]"""
    full = np.convolve(x, f)            # full[n] = sum_j f[j] x[n - j]
    return full[1::2][: len(x) // 2]


def test_haar_oracle_matches_filter_bank_definition():
    from oracle import haar
    s = 1 / np.sqrt(2.0)
    dec_lo, dec_hi = np.array([s, s]), np.array([-s, s])           # pywt.Wavelet('haar').dec_lo / dec_hi (documented)
    rng = np.random.default_rng(0)
    x = rng.standard_normal(16)
    lo, hi = haar.dwt(x)
    assert np.abs(lo - _bank(x, dec_lo)).max() < 1e-12 and np.abs(hi - _bank(x, dec_hi)).max() < 1e-12
    assert np.allclose(_bank(np.array([1.0, 2, 3, 4]), dec_hi), [-s, -s])       # the documented pywt KAT
    img = rng.standard_normal((8, 12))
    cA, (cH, cV, cD) = haar.dwt2(img)
    rows_lo = np.stack([_bank(c, dec_lo) for c in img.T], 1)       # along axis 0 first ...
    rows_hi = np.stack([_bank(c, dec_hi) for c in img.T], 1)

    def along1(m, f):
        return np.stack([_bank(r, f) for r in m], 0)

    # ... then axis 1; dwt2 returns (aa, (da, ad, dd)), first letter = axis 0
    for got, want in ((cA, along1(rows_lo, dec_lo)), (cH, along1(rows_hi, dec_lo)), (cV, along1(rows_lo, dec_hi)),
                      (cD, along1(rows_hi, dec_hi))):
        assert np.abs(got - want).max() < 1e-12
    back = haar.idwt2((cA, (cH, cV, cD)))
    assert np.abs(back - img).max() < 1e-12
    # 2x2 block closed form of SURVEY App. A.3 on random blocks
    a, b, c, d = (img[0::2, 0::2], img[0::2, 1::2], img[1::2, 0::2], img[1::2, 1::2])
    assert np.abs(cA - (a + b + c + d) / 2).max() < 1e-12 and np.abs(cH - (a + b - c - d) / 2).max() < 1e-12
    assert np.abs(cV - (a - b + c - d) / 2).max() < 1e-12 and np.abs(cD - (a - b - c + d) / 2).max() < 1e-12


def test_attention_oracle_equals_explicit_softmax():
    """the Attention restatement (attention_processor.AttnProcessor2_0 path) against the textbook formula, incl. the
    single-token cross-attention identity the product's collapsed attn2 relies on (softmax over one key == 1)."""
    from oracle import diffusers_restated as dr
    torch.manual_seed(0)
    att = dr.Attention(64, heads=4, dim_head=16, bias=False).double()
    x = torch.randn(2, 9, 64, dtype=torch.float64)
    q, k, v = att.to_q(x), att.to_k(x), att.to_v(x)
    sp = lambda t: t.reshape(2, 9, 4, 16).transpose(1, 2)
    p = torch.softmax(sp(q) @ sp(k).transpose(-1, -2) / 4.0, -1)
    want = att.to_out[0]((p @ sp(v)).transpose(1, 2).reshape(2, 9, 64))
    assert (att(x) - want).abs().max() < 1e-12
    xatt = dr.Attention(64, heads=4, dim_head=16, cross_attention_dim=24).double()
    ctx = torch.randn(2, 1, 24, dtype=torch.float64)
    want1 = xatt.to_out[0](xatt.to_v(ctx)).expand(2, 9, 64)
    assert (xatt(x, ctx) - want1).abs().max() < 1e-12


# ---- independent third-party implementations of diffusers' blocks that ship inside `transformers` -------------------------------
# The VQ-VAE encoders of Chameleon / Emu3 carry the latent-diffusion ResnetBlock and Upsample that diffusers' ResnetBlock2D
# and Upsample2D descend from, and Qwen2.5-Omni's token2wav DiT carries AdaLayerNormZero, the adaLN-Zero decoder layer
# and the GELU(tanh) MLP of diffusers' BasicTransformerBlock(norm_type="ada_norm_zero").  They were written by other
# people from the same upstream sources: executed where they lie, with the oracle's weights copied in.
def _tf(modname, cls):
    try:
        with _without_test_stubs():
            mod = __import__("transformers.models." + modname, fromlist=[cls])
        return getattr(mod, cls)
    except Exception as e:  # pragma: no cover
        pytest.skip(f"transformers.{modname}.{cls} not importable: {e}")


class _Const(torch.nn.Module):
    """stands in for a sub-module whose output the test wants to dictate"""

    def __init__(self, value):
        super().__init__()
        self.value = value

    def forward(self, *args, **kwargs):
        return self.value


@pytest.mark.parametrize("cin,cout", [(64, 64), (64, 128)])
def test_resnet_block_matches_transformers_vqvae_resnet_block(cin, cout):
    """oracle ResnetBlock2D (GroupNorm -> SiLU -> conv3x3 -> [+ time embedding] -> GroupNorm -> SiLU -> conv3x3, 1x1 shortcut
    when the channel count changes, residual add) against ChameleonVQVAEEncoderResnetBlock with the time-embedding
    projection zeroed (the VQ-VAE block has none)."""
    from types import SimpleNamespace
    from oracle import diffusers_restated as dr
    Blk = _tf("chameleon.modeling_chameleon", "ChameleonVQVAEEncoderResnetBlock")
    torch.manual_seed(cin + cout)
    ours = dr.ResnetBlock2D(cin, cout, temb_channels=32, groups=32, eps=1e-6).double()
    with torch.no_grad():
        ours.time_emb_proj.weight.zero_()
        ours.time_emb_proj.bias.zero_()
        for p in (ours.norm1.weight, ours.norm1.bias, ours.norm2.weight, ours.norm2.bias):
            p.copy_(torch.randn_like(p))
    ref = Blk(SimpleNamespace(dropout=0.0), cin, cout).double().eval()
    pairs = [("norm1", "norm1"), ("conv1", "conv1"), ("norm2", "norm2"), ("conv2", "conv2")]
    if cin != cout:
        pairs.append(("nin_shortcut", "conv_shortcut"))
    for theirs, mine in pairs:
        getattr(ref, theirs).load_state_dict(getattr(ours, mine).state_dict())
    x = torch.randn(2, cin, 8, 8, dtype=torch.float64)
    temb = torch.randn(2, 32, dtype=torch.float64)
    assert (ours(x, temb) - ref(x.clone())).abs().max() < 1e-10


def test_upsample_matches_transformers_vqvae_upsample():
    """oracle Upsample2D (nearest 2x, then conv3x3 pad 1) against Emu3VQVAEEncoderConvUpsample."""
    from oracle import diffusers_restated as dr
    Up = _tf("emu3.modeling_emu3", "Emu3VQVAEEncoderConvUpsample")
    torch.manual_seed(3)
    ours, ref = dr.Upsample2D(16).double(), Up(16).double()
    ref.conv.load_state_dict(ours.conv.state_dict())
    x = torch.randn(2, 16, 5, 7, dtype=torch.float64)
    assert (ours(x) - ref(x)).abs().max() < 1e-12


def test_adaln_zero_block_matches_transformers_dit():
    """oracle AdaLayerNormZero (SiLU -> Linear(6 dim) -> chunks IN THE ORDER shift_msa, scale_msa, gate_msa, shift_mlp,
    scale_mlp, gate_mlp; LayerNorm without affine, eps 1e-6; x (1 + scale) + shift) and the feed-forward half of the adaLN-Zero
    block (h + gate_mlp * MLP(LN(h) (1 + scale_mlp) + shift_mlp), MLP = Linear -> GELU(tanh) -> Linear) against
    Qwen2.5-Omni's DiT (Qwen2_5_OmniAdaLayerNormZero, DiTMLP and the arithmetic of its decoder layer)."""
    from oracle import diffusers_restated as dr
    Ada = _tf("qwen2_5_omni.modeling_qwen2_5_omni", "Qwen2_5_OmniAdaLayerNormZero")
    MLP = _tf("qwen2_5_omni.modeling_qwen2_5_omni", "DiTMLP")
    torch.manual_seed(5)
    dim, B, N = 48, 3, 7
    blk = dr.BasicTransformerBlockAdaZero(dim, heads=4, dim_head=12, num_embeds=10, norm_eps=1e-6, attention_bias=True).double()
    emb = torch.randn(B, dim, dtype=torch.float64)
    blk.norm1.emb = _Const(emb)                                  # feed the conditioning vector directly
    ref_ada = Ada(dim).double()
    ref_ada.linear.load_state_dict(blk.norm1.linear.state_dict())
    x = torch.randn(B, N, dim, dtype=torch.float64)
    got = blk.norm1(x, None, None)
    want = ref_ada(x, emb=emb)
    for g, w in zip(got, want):
        assert (g - w).abs().max() < 1e-12
    ref_mlp = MLP(dim, mult=4).double()
    ref_mlp.ff[0].load_state_dict(blk.ff.net[0].proj.state_dict())
    ref_mlp.ff[3].load_state_dict(blk.ff.net[2].state_dict())
    # the block with its attention output replaced by a known tensor: Qwen's decoder-layer arithmetic
    a_out = torch.randn(B, N, dim, dtype=torch.float64)
    blk.attn1 = _Const(a_out)
    _, gate_msa, shift_mlp, scale_mlp, gate_mlp = want
    h = x + gate_msa.unsqueeze(1) * a_out
    norm = torch.nn.functional.layer_norm(h, (dim,), eps=1e-6) * (1 + scale_mlp[:, None]) + shift_mlp[:, None]
    ref_out = h + gate_mlp.unsqueeze(1) * ref_mlp(norm)
    assert (blk(x, None, None) - ref_out).abs().max() < 1e-11


# ---- published parameter counts of well-known checkpoints: pin the WIRING of the restated models (every channel count of the
# skip concatenations, the transformer blocks' inner sizes, the conditioning MLPs) independently of the reference's own configs
def test_unet_restatement_reproduces_stable_diffusion_v1_5_parameter_count():
    """diffusers' UNet2DConditionModel at the Stable Diffusion v1.x config (block_out_channels 320/640/1280/1280, 2 layers
    per block, 8 heads, cross_attention_dim 768, CrossAttn x3 + Down / Up + CrossAttnUp x3) has 859 520 964 parameters
    (the number every SD 1.x model card and `diffusers` issue quotes).  The oracle class, built on the meta device with
    that config, must have exactly that many once the reference-specific ``encoder_hid_proj`` is taken out."""
    from oracle import diffusers_restated as dr
    with torch.device("meta"):
        m = dr.UNet2DConditionModel(sample_size=64, in_channels=4, out_channels=4, block_out_channels=(320, 640, 1280, 1280),
                                    layers_per_block=2, cross_attention_dim=768, attention_head_dim=8, encoder_hid_dim=16,
                                    encoder_hid_dim_type="text_proj")
    n = sum(p.numel() for p in m.parameters()) - sum(p.numel() for p in m.encoder_hid_proj.parameters())
    assert n == 859_520_964


def test_dit_restatement_reproduces_dit_xl_2_parameter_count():
    """facebookresearch/DiT prints ``DiT Parameters: 675,129,632`` for DiT-XL/2 (256 x 256: 32 x 32 x 4 latents, patch 2,
    28 layers, 16 heads x 72, learn_sigma, 1000 classes + the null class).  That count includes the frozen sincos table
    (256 x 1152, an nn.Parameter there, a buffer in diffusers); diffusers' port additionally gives EVERY block's
    AdaLayerNormZero its own copy of the timestep / label embedder (28 instead of 1).  Accounting for exactly those two
    differences the oracle's DiTTransformer2DModel has the published count."""
    from oracle import diffusers_restated as dr
    with torch.device("meta"):
        m = dr.DiTTransformer2DModel(num_attention_heads=16, attention_head_dim=72, in_channels=4, out_channels=8, num_layers=28,
                                     sample_size=32, patch_size=2, num_embeds_ada_norm=1000, attention_bias=True,
                                     norm_num_groups=32)
    total = sum(p.numel() for p in m.parameters())
    embedder = sum(p.numel() for p in m.transformer_blocks[0].norm1.emb.parameters())
    assert total - 27 * embedder + 256 * 1152 == 675_129_632


def test_product_unet_class_has_the_same_published_parameter_count():
    """dcb200.UNetCondition2D keeps diffusers' parameter names and shapes (reference checkpoints load into it): at the
    Stable Diffusion v1.x config it too has 859 520 964 parameters + encoder_hid_proj."""
    import dcb200
    with torch.device("meta"):
        m = dcb200.UNetCondition2D(sample_size=64, in_channels=4, out_channels=4, block_out_channels=(320, 640, 1280, 1280),
                                   layers_per_block=2, cross_attention_dim=768, attention_head_dim=8, encoder_hid_dim=16,
                                   encoder_hid_dim_type="text_proj",
                                   down_block_types=("CrossAttnDownBlock2D",) * 3 + ("DownBlock2D",),
                                   up_block_types=("UpBlock2D",) + ("CrossAttnUpBlock2D",) * 3)
    assert sum(p.numel() for p in m.parameters()) - (16 * 768 + 768) == 859_520_964


def test_geglu_feed_forward_matches_transformers_clvp_gated_mlp():
    """oracle GEGLU / FeedForward("geglu") -- one Linear to 2 x inner, FIRST half the value, SECOND half the gate,
    value * gelu(gate) with the exact (erf) GELU, then Linear(inner, dim) -- against CLVP's ClvpGatedLinearUnit / ClvpEncoderMLP
    (tortoise-tts' port of the same x-transformers / latent-diffusion GEGLU) with hidden_act = "gelu"."""
    from types import SimpleNamespace
    from oracle import diffusers_restated as dr
    MLP = _tf("clvp.modeling_clvp", "ClvpEncoderMLP")
    torch.manual_seed(11)
    dim = 24
    ours = dr.FeedForward(dim, "geglu").double()
    ref = MLP(SimpleNamespace(hidden_act="gelu", hidden_size=dim, intermediate_size=4 * dim, dropout=0.0)).double().eval()
    ref.fc1.proj.load_state_dict(ours.net[0].proj.state_dict())
    ref.fc2.load_state_dict(ours.net[2].state_dict())
    x = torch.randn(3, 5, dim, dtype=torch.float64)
    assert (ours(x) - ref(x)).abs().max() < 1e-12
    assert (ours.net[0](x) - ref.fc1(x)).abs().max() < 1e-12
