"""drop-in for /root/reference/utils/metrics.py (counters stay on the GPU; no per-batch host sync)."""
from dcb200.metrics import F1, Accuracy, Metric, Precision, Recall  # noqa: F401
