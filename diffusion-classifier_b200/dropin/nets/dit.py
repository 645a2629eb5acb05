"""drop-in for /root/reference/nets/dit.py:8-51."""
from dcb200.dit import DiT  # noqa: F401
