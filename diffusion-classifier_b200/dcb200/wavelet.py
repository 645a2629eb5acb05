"""Haar wavelet transforms with the reference's signatures (utils/wavelet.py:4-35, 37-67), on the GPU.

``wavelet_dec_2(images[C,H,W]) -> [4C,H/2,W/2]`` and ``wavelet_enc_2(w[4C,h,w]) -> [C,2h,2w]`` (channel order
4i+{0,1,2,3} = cA,cH,cV,cD); batched [B,C,H,W] inputs are accepted too.  CPU tensors are moved to the current
CUDA device and back (the reference runs pywt on the CPU inside DataLoader workers)."""
import torch

from . import engine as E


def _run(fn, t, scale):
    squeeze = t.dim() == 3
    src_dev = t.device
    if squeeze:
        t = t.unsqueeze(0)
    if not t.is_cuda:
        if not torch.cuda.is_available():
            raise RuntimeError("dcb200.wavelet needs a CUDA device; there is no CPU path")
        t = t.cuda()
    out = fn(t, scale)
    if squeeze:
        out = out[0]
    return out.to(src_dev)


def wavelet_dec_2(images, post_scale=1.0):
    return _run(E.haar_dwt, images, post_scale)


def wavelet_enc_2(wavelet_images, pre_scale=1.0):
    return _run(E.haar_idwt, wavelet_images, pre_scale)
