"""Shared test helpers: model/config builders and fp32 torch restatements of individual ops (checkers only)."""
import torch
import torch.nn.functional as F

TINY_UNET = dict(
    sample_size=16, in_channels=3, out_channels=3, layers_per_block=1, block_out_channels=(64, 256),
    down_block_types=("DownBlock2D", "CrossAttnDownBlock2D"), up_block_types=("CrossAttnUpBlock2D", "UpBlock2D"),
    mid_block_type="UNetMidBlock2DCrossAttn", encoder_hid_dim=64, encoder_hid_dim_type="text_proj",
    cross_attention_dim=64)

# a 3-level net that exercises every structural feature of the BASELINE configs at small cost: tuple
# layers_per_block, DownBlock->CrossAttnDown->Down, CrossAttnUp, concat channel counts that straddle GN groups
SMALL_UNET = dict(
    sample_size=32, in_channels=12, out_channels=12, layers_per_block=(1, 2, 1), block_out_channels=(64, 256, 512),
    down_block_types=("DownBlock2D", "CrossAttnDownBlock2D", "DownBlock2D"),
    up_block_types=("UpBlock2D", "CrossAttnUpBlock2D", "UpBlock2D"), mid_block_type="UNetMidBlock2DCrossAttn",
    encoder_hid_dim=128, encoder_hid_dim_type="text_proj", cross_attention_dim=128)

# the BASELINE.json architectures live in the product (bench.py builds its workloads from them)
from dcb200.configs import CIFAR_UNET, DIT_B4_256, IPMSA5_DWT_UNET, UNET128, UNET256  # noqa: E402,F401

TINY_DIT = dict(num_attention_heads=2, attention_head_dim=64, in_channels=3, out_channels=3, num_layers=2,
                sample_size=32, patch_size=2, norm_eps=1e-5)


from dcb200.configs import Config as Cfg  # noqa: E402  (duck-typed config like the experiments')
from dcb200.configs import classify_config as base_cfg  # noqa: E402,F401


def rel_err(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def max_rel(a, b):
    a, b = a.float(), b.float()
    return float(((a - b).abs() / b.abs().clamp_min(1e-6)).max())


def make_pair(kind, arch, seed=0, device="cpu", amplify=1.0):
    """(oracle module, dcb200 module) with identical weights (default torch init under manual_seed)."""
    from oracle import diffusers_restated as dr
    import dcb200
    torch.manual_seed(seed)
    if kind == "unet":
        o = dr.UNet2DConditionModel(**arch)
        p = dcb200.UNetCondition2D(**arch)
    else:
        o = dr.DiTTransformer2DModel(**arch)
        p = dcb200.DiT(**arch)
    if amplify != 1.0 and kind == "dit":  # random-init nets barely react to the class: make margins meaningful
        with torch.no_grad():
            for b in o.transformer_blocks:
                b.norm1.emb.class_embedder.embedding_table.weight.mul_(amplify)
    p.load_state_dict(o.state_dict())
    return o.to(device).eval(), p.to(device).eval()


def conv_ref(x_nhwc, w_oihw, bias, stride=1, pad=1):
    """x [N,H,W,C] fp32 -> NHWC conv via torch."""
    y = F.conv2d(x_nhwc.permute(0, 3, 1, 2), w_oihw, bias, stride=stride, padding=pad)
    return y.permute(0, 2, 3, 1).contiguous()
