"""Architectures of the BASELINE.json configs as keyword dictionaries for ``UNetCondition2D`` / ``DiT`` -- the values the
reference's model fragments pass (``config.image_size`` etc. already substituted):

    UNET128          models/unet-128.py:4-27          (image_size 128, 3 channels, no wavelet transform)
    UNET256          models/unet-256.py:4-30          (image_size 256)
    DIT_B4_256       models/chexpert-256-dit-b4.py:4-21   (patch_size 4)
    IPMSA5_DWT_UNET  models/ipmsa-5-dwt-unet.py:4-28  (10 pixel channels, wavelet transform: 40 x 128 x 128)
    CIFAR_UNET       experiments/cifar10/inference.py:94-116

``tests/test_cpu_dropin.py`` executes the reference's own fragments against ``dropin/`` and asserts that they build
exactly these networks; ``bench.py`` builds its workloads from here."""

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

CIFAR_UNET = dict(
    sample_size=32, in_channels=3, out_channels=3, layers_per_block=2, block_out_channels=(128, 128, 256, 512),
    down_block_types=("DownBlock2D", "DownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D"),
    up_block_types=("CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "UpBlock2D", "UpBlock2D"),
    mid_block_type="UNetMidBlock2DCrossAttn", encoder_hid_dim=128, encoder_hid_dim_type="text_proj",
    cross_attention_dim=128)


class Config:
    """Attribute view of a plain dict, like the ``TrainingConfig`` of the reference's experiment scripts
    (experiments/chexpert-unet/inference.py:24-39): missing keys read ``None``."""

    def __init__(self, **kw):
        self.__dict__["_d"] = dict(kw)

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return self.__dict__["_d"].get(name)

    def __setattr__(self, name, value):
        self.__dict__["_d"][name] = value


def classify_config(**kw):
    """the keys ``DiffusionClassifier`` reads on the classification path, with the experiments' usual values"""
    d = dict(pred_param="eps", schedule="cosine", noise_d=32, image_size=32, cfg_w=0.0, ema_beta=0.999, ema_warmup=0,
             ema_update_freq=1, encoder_type="nn", classes=4, n_stages=1, evaluation_per_stage=[4],
             n_keep_per_stage=[1], n_fast_classes=2, fast_classification=False)
    d.update(kw)
    return Config(**d)
