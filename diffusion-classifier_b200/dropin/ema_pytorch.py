"""stand-in for the ema-pytorch package where it is not installed (diffusion_classifier.py:10)."""
from dcb200.ema import EMA  # noqa: F401
