"""drop-in for /root/reference/diffusion/diffusion_classifier.py: classify (hot path), sample, loss (forward),
evaluate / inference, load_checkpoint.  train_loop (backward, optimiser) is outside this library."""
from dcb200.classifier import DiffusionClassifier, log  # noqa: F401
