"""drop-in for /root/reference/diffusion/diffusion_classifier.py (classification path only)."""
from dcb200.classifier import DiffusionClassifier, log  # noqa: F401
