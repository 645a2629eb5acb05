"""ORACLE (test infrastructure only).  numpy restatement of ``pywt.dwt2 / pywt.idwt2(..., 'haar')``.

pywt (PyWavelets, unpinned: absent from /root/reference/requirements.txt, not installed here) is what
the reference calls at utils/wavelet.py:27 (dwt2) and utils/wavelet.py:63 (idwt2).  Published algorithm
for the Haar filter bank with even-length signals (no boundary extension is exercised):
    lo[k] = (x[2k] + x[2k+1]) / sqrt(2),   hi[k] = (x[2k] - x[2k+1]) / sqrt(2)
``dwt2`` applies it along axis 0 then axis 1 and returns ``cA, (cH, cV, cD)`` = (aa, (da, ad, dd)) where the
first letter refers to axis 0 (rows).  Documented KAT: pywt.dwt([1,2,3,4],'haar') = ([2.1213, 4.9497],
[-0.7071, -0.7071]).  PARITY UNPINNED beyond that KAT (see oracle/__init__.py).
"""
import sys
import types

import numpy as np

_S = np.sqrt(2.0)


def dwt(x, wavelet="haar", axis=-1):
    assert wavelet == "haar"
    x = np.moveaxis(np.asarray(x, dtype=np.float64), axis, -1)
    assert x.shape[-1] % 2 == 0, "only even lengths are used by the reference configs"
    lo = (x[..., 0::2] + x[..., 1::2]) / _S
    hi = (x[..., 0::2] - x[..., 1::2]) / _S
    return np.moveaxis(lo, -1, axis), np.moveaxis(hi, -1, axis)


def idwt(lo, hi, wavelet="haar", axis=-1):
    assert wavelet == "haar"
    lo = np.moveaxis(np.asarray(lo, dtype=np.float64), axis, -1)
    hi = np.moveaxis(np.asarray(hi, dtype=np.float64), axis, -1)
    out = np.empty(lo.shape[:-1] + (2 * lo.shape[-1],), dtype=np.float64)
    out[..., 0::2] = (lo + hi) / _S
    out[..., 1::2] = (lo - hi) / _S
    return np.moveaxis(out, -1, axis)


def dwt2(data, wavelet="haar"):
    data = np.asarray(data)
    dt = data.dtype if data.dtype in (np.float32, np.float64) else np.float64
    a, d = dwt(data, wavelet, axis=0)          # along rows (axis 0)
    aa, ad = dwt(a, wavelet, axis=1)
    da, dd = dwt(d, wavelet, axis=1)
    return aa.astype(dt), (da.astype(dt), ad.astype(dt), dd.astype(dt))


def idwt2(coeffs, wavelet="haar"):
    aa, (da, ad, dd) = coeffs
    dt = np.asarray(aa).dtype
    a = idwt(aa, ad, wavelet, axis=1)
    d = idwt(da, dd, wavelet, axis=1)
    return idwt(a, d, wavelet, axis=0).astype(dt if dt in (np.float32, np.float64) else np.float64)


def wavelet_dec_2_np(images):
    """[C,H,W] -> [4C,H/2,W/2], channel order 4i+{0,1,2,3} = cA,cH,cV,cD (utils/wavelet.py:4-35)."""
    C, H, W = images.shape
    out = np.zeros((4 * C, H // 2, W // 2), dtype=np.float32)
    for i in range(C):
        cA, (cH, cV, cD) = dwt2(images[i])
        out[4 * i], out[4 * i + 1], out[4 * i + 2], out[4 * i + 3] = cA, cH, cV, cD
    return out


def wavelet_enc_2_np(w):
    """[4C,h,w] -> [C,2h,2w] (utils/wavelet.py:37-67)."""
    C = w.shape[0] // 4
    out = np.zeros((C, w.shape[1] * 2, w.shape[2] * 2), dtype=np.float32)
    for i in range(C):
        out[i] = idwt2((w[4 * i], (w[4 * i + 1], w[4 * i + 2], w[4 * i + 3])))
    return out


def as_pywt_module():
    """A stand-in ``pywt`` module so the reference's utils/wavelet.py can be imported verbatim."""
    m = types.ModuleType("pywt")
    m.dwt, m.idwt, m.dwt2, m.idwt2 = dwt, idwt, dwt2, idwt2
    return m
