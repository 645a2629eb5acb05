"""drop-in for /root/reference/nets/unet.py:77-195."""
from dcb200.unet import UNetCondition2D  # noqa: F401
