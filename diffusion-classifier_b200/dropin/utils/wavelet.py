"""drop-in for /root/reference/utils/wavelet.py (Haar DWT / IDWT on the GPU)."""
from dcb200.wavelet import wavelet_dec_2, wavelet_enc_2  # noqa: F401
