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

CIFAR_UNET = dict(
    sample_size=32, in_channels=3, out_channels=3, layers_per_block=2, block_out_channels=(128, 128, 256, 512),
    down_block_types=("DownBlock2D", "DownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D"),
    up_block_types=("CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "UpBlock2D", "UpBlock2D"),
    mid_block_type="UNetMidBlock2DCrossAttn", encoder_hid_dim=128, encoder_hid_dim_type="text_proj",
    cross_attention_dim=128)

UNET128 = dict(
    sample_size=128, in_channels=3, out_channels=3, layers_per_block=2, block_out_channels=(128, 128, 256, 512, 1024),
    down_block_types=("DownBlock2D", "DownBlock2D", "DownBlock2D", "CrossAttnDownBlock2D", "DownBlock2D"),
    up_block_types=("UpBlock2D", "CrossAttnUpBlock2D", "UpBlock2D", "UpBlock2D", "UpBlock2D"),
    mid_block_type="UNetMidBlock2DCrossAttn", encoder_hid_dim=512, encoder_hid_dim_type="text_proj",
    cross_attention_dim=512)

UNET256 = dict(
    sample_size=256, in_channels=3, out_channels=3, layers_per_block=2,
    block_out_channels=(128, 128, 256, 256, 512, 1024),
    down_block_types=("DownBlock2D", "DownBlock2D", "DownBlock2D", "DownBlock2D", "CrossAttnDownBlock2D", "DownBlock2D"),
    up_block_types=("UpBlock2D", "CrossAttnUpBlock2D", "UpBlock2D", "UpBlock2D", "UpBlock2D", "UpBlock2D"),
    mid_block_type="UNetMidBlock2DCrossAttn", encoder_hid_dim=512, encoder_hid_dim_type="text_proj",
    cross_attention_dim=512)

IPMSA5_DWT_UNET = dict(
    sample_size=128, in_channels=40, out_channels=40, layers_per_block=(2, 2, 2, 4, 2),
    block_out_channels=(128, 128, 256, 512, 768),
    down_block_types=("DownBlock2D", "DownBlock2D", "DownBlock2D", "CrossAttnDownBlock2D", "DownBlock2D"),
    up_block_types=("UpBlock2D", "CrossAttnUpBlock2D", "UpBlock2D", "UpBlock2D", "UpBlock2D"),
    mid_block_type="UNetMidBlock2DCrossAttn", encoder_hid_dim=512, encoder_hid_dim_type="text_proj",
    cross_attention_dim=512)

DIT_B4_256 = dict(num_attention_heads=12, attention_head_dim=64, in_channels=3, out_channels=3, num_layers=12,
                  dropout=0.0, norm_num_groups=32, attention_bias=True, sample_size=256, patch_size=4,
                  activation_fn="gelu-approximate", num_embeds_ada_norm=1000, upcast_attention=False,
                  norm_type="ada_norm_zero", norm_elementwise_affine=False, norm_eps=1e-5)

TINY_DIT = dict(num_attention_heads=2, attention_head_dim=64, in_channels=3, out_channels=3, num_layers=2,
                sample_size=32, patch_size=2, norm_eps=1e-5)


class Cfg:
    """Duck-typed config like the experiments' (missing keys read None; dunders raise so deepcopy works)."""

    def __init__(self, **kw):
        self.__dict__["_d"] = dict(kw)

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return self.__dict__["_d"].get(name)

    def __setattr__(self, name, value):
        self.__dict__["_d"][name] = value


def base_cfg(**kw):
    d = dict(pred_param="eps", schedule="cosine", noise_d=32, image_size=32, cfg_w=0.0, ema_beta=0.999, ema_warmup=0,
             ema_update_freq=1, encoder_type="nn", classes=4, n_stages=1, evaluation_per_stage=[4],
             n_keep_per_stage=[1], n_fast_classes=2, fast_classification=False)
    d.update(kw)
    return Cfg(**d)


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
