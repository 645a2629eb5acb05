"""dcb200 -- B200-native (sm_100a) implementation of the classification-by-ELBO hot path of
faverogian/diffusion-classifier, behind the reference's own Python call signatures.

    from dcb200 import DiffusionClassifier, UNetCondition2D, DiT, EMA, wavelet_dec_2, wavelet_enc_2

All arithmetic runs in libdcb200.so (hand-written CUDA, C ABI in include/dcb200.h); there is no CPU or torch
fallback -- importing the ops without the built library raises.
"""
from ._lib import DcbError, launch_count, lib  # noqa: F401
from .classifier import DiffusionClassifier  # noqa: F401
from .dit import DiT  # noqa: F401
from . import metrics  # noqa: F401
from .ema import EMA  # noqa: F401
from .unet import UNetCondition2D  # noqa: F401
from .wavelet import wavelet_dec_2, wavelet_enc_2  # noqa: F401

__version__ = "0.1.0"
