"""BASELINE.json configs C2-C5 at their REAL architectures and image sizes (models/unet-128.py, models/unet-256.py,
models/chexpert-256-dit-b4.py, models/ipmsa-5-dwt-unet.py): product classify() on the B200 vs the fp32 oracle loop on
the host with identical pre-drawn noise, few timesteps (the oracle must finish in seconds).  bf16 tolerance 1e-2 on the
per-class ELBO errors (north star); margin-aware label comparison."""
import numpy as np
import pytest
import torch

from helpers import DIT_B4_256, IPMSA5_DWT_UNET, UNET128, UNET256, base_cfg, make_pair
from test_gpu_c_models import _check_classify

pytestmark = pytest.mark.gpu


def _run(dev, kind, arch, cfg, x, T, BS, amplify):
    import dcb200
    from oracle import loop
    o, p = make_pair(kind, arch, seed=0, amplify=amplify)
    torch.manual_seed(1)
    dc = dcb200.DiffusionClassifier(p, cfg)
    enc = None
    if kind == "unet":
        with torch.no_grad():
            dc.encoder.weight.mul_(amplify)
        enc = torch.nn.Embedding(cfg.classes + 1, arch["encoder_hid_dim"])
        enc.load_state_dict(dc.encoder.state_dict())
    g = torch.Generator().manual_seed(2)
    t_all = torch.rand(T, BS, generator=g)
    eps_all = torch.randn(T, *x.shape, generator=g)

    class Den(torch.nn.Module):
        def forward(self, x, noise_labels, encoder_hidden_states):
            return o(x, noise_labels, encoder_hidden_states)[0]

    ref_labels, ref_err = loop.classify_oracle(Den(), enc, cfg, x, t_all=t_all, eps_all=eps_all, return_errors=True)
    dc = dc.to(dev).eval()
    labels = dc.classify(x.to(dev), t_all=t_all, eps_all=eps_all)
    _check_classify(labels.cpu(), dc.last_errors, ref_labels.numpy(), ref_err.mean(dim=2).numpy(), 1e-2)
    if kind == "unet":   # shared class-independent prefix (default) vs the reference's per-class recomputation
        e1 = dc.last_errors.clone()
        cfg.dcb_share_prefix = False
        dc.classify(x.to(dev), t_all=t_all, eps_all=eps_all)
        assert torch.equal(dc.last_errors, e1)
        cfg.dcb_share_prefix = None
    return dc


def test_c2_unet128(dev):
    cfg = base_cfg(classes=2, evaluation_per_stage=[2], noise_d=128, image_size=128)
    x = torch.rand(2, 3, 128, 128, generator=torch.Generator().manual_seed(0)) * 2 - 1
    _run(dev, "unet", UNET128, cfg, x, 2, 2, 40.0)


def test_c3_unet256_shifted_cosine(dev):
    cfg = base_cfg(classes=2, evaluation_per_stage=[1], noise_d=64, image_size=256, schedule="shifted_cosine")
    x = torch.rand(1, 3, 256, 256, generator=torch.Generator().manual_seed(0)) * 2 - 1
    _run(dev, "unet", UNET256, cfg, x, 1, 1, 40.0)


def test_c4_dit_b4_256(dev):
    cfg = base_cfg(classes=2, evaluation_per_stage=[1], noise_d=64, image_size=256, schedule="shifted_cosine",
                   encoder_type="DiT", pred_param="v")
    x = torch.rand(1, 3, 256, 256, generator=torch.Generator().manual_seed(0)) * 2 - 1
    _run(dev, "dit", DIT_B4_256, cfg, x, 1, 1, 20.0)


def test_c5_ipmsa5_dwt_unet_with_gpu_haar(dev):
    """pixel input [BS,10,256,256] -> Haar DWT /2 (utils/wavelet.py + experiments/ipmsa/inference.py:153-155) ->
    [BS,40,128,128] wavelet-domain ELBO classification; DWT by dcb_haar_dwt vs the numpy oracle."""
    import dcb200
    from oracle import haar
    pix = torch.rand(1, 10, 256, 256, generator=torch.Generator().manual_seed(0)) * 2 - 1
    w_ref = torch.from_numpy(np.stack([haar.wavelet_dec_2_np(im.numpy()) for im in pix])) / 2
    w = dcb200.wavelet_dec_2(pix.to(dev), 0.5)
    assert w.shape == (1, 40, 128, 128) and (w.cpu() - w_ref).abs().max() < 1e-6
    cfg = base_cfg(classes=2, evaluation_per_stage=[1], noise_d=128, image_size=128, wavelet_transform=True)
    _run(dev, "unet", IPMSA5_DWT_UNET, cfg, w_ref, 1, 1, 40.0)


# ---- size-independent properties at the BASELINE configs' full tensor sizes ------------------------------------------------
def test_c5_haar_roundtrip_and_energy_at_full_size(dev):
    """IDWT(DWT(x)) == x and Parseval (orthonormal Haar: sum of squares preserved) on a full IPMSA batch [8,10,256,256]."""
    import dcb200
    x = torch.rand(8, 10, 256, 256, device=dev, generator=torch.Generator(device=dev).manual_seed(0)) * 2 - 1
    w = dcb200.wavelet_dec_2(x)
    assert w.shape == (8, 40, 128, 128)
    assert abs(float((w.double() ** 2).sum()) / float((x.double() ** 2).sum()) - 1) < 1e-6
    back = dcb200.wavelet_enc_2(w)
    assert (back - x).abs().max() < 1e-6
    # linearity of the transform and the /2 scaling used by experiments/ipmsa/inference.py:153-155
    y = torch.rand_like(x)
    assert (dcb200.wavelet_dec_2(x + 2 * y) - (w + 2 * dcb200.wavelet_dec_2(y))).abs().max() < 1e-5
    assert torch.equal(dcb200.wavelet_dec_2(x, 0.5), w * 0.5)


@pytest.mark.parametrize("arch,C,S", [(UNET128, 3, 128), (IPMSA5_DWT_UNET, 40, 128)], ids=["c2_unet128", "c5_ipmsa5"])
def test_unet_error_table_invariant_to_chunking_and_graphs(dev, arch, C, S):
    """unet-128 / ipmsa-5-dwt at their real sizes: the per-(image, class, timestep) error table is bit-identical whatever the launch-sequence
    size (dcb_max_batch: different tile counts select different tcgen05 kernels / halo modes), with CUDA-graph replay or
    eager launches, and with or without the shared class-independent prefix; in-kernel Philox noise is a function of
    (seed, unit) only."""
    import dcb200
    torch.manual_seed(0)
    net = dcb200.UNetCondition2D(**arch)
    cfg = base_cfg(classes=2, evaluation_per_stage=[6], noise_d=S, image_size=S)
    dc = dcb200.DiffusionClassifier(net, cfg).to(dev).eval()
    x = torch.rand(3, C, S, S, device=dev, generator=torch.Generator(device=dev).manual_seed(1)) * 2 - 1
    tables = []
    for mb, graph, share in ((0, None, None), (8, None, None), (14, False, None), (36, None, False), (2, False, False)):
        cfg.dcb_max_batch, cfg.dcb_cuda_graph, cfg.dcb_share_prefix = mb, graph, share
        for _ in range(3 if graph is None else 1):     # third call replays the captured graph
            dc._eps_calls = 0
            torch.manual_seed(5)
            labels = dc.classify(x)
        tables.append((dc.last_errors.clone(), labels.clone()))
    for t, l in tables[1:]:
        assert torch.equal(t, tables[0][0]) and torch.equal(l, tables[0][1])
    assert torch.isfinite(tables[0][0]).all()


def test_c4_dit_b4_error_table_invariant_to_chunking(dev):
    """DiT-B/4 256 at its real size: bit-identical error table for every launch-sequence size (tile counts select
    different tile widths / epilogues / tcgen05 kernels; attention picks its kernel per (batch, head))."""
    import dcb200
    torch.manual_seed(0)
    net = dcb200.DiT(**DIT_B4_256)
    cfg = base_cfg(classes=2, evaluation_per_stage=[3], noise_d=64, image_size=256, schedule="shifted_cosine",
                   encoder_type="DiT", pred_param="v")
    dc = dcb200.DiffusionClassifier(net, cfg).to(dev).eval()
    x = torch.rand(2, 3, 256, 256, device=dev, generator=torch.Generator(device=dev).manual_seed(1)) * 2 - 1
    tables = []
    for mb, graph in ((0, False), (2, False), (6, None), (4, False)):
        cfg.dcb_max_batch, cfg.dcb_cuda_graph = mb, graph
        for _ in range(3 if graph is None else 1):
            dc._eps_calls = 0
            torch.manual_seed(5)
            dc.classify(x)
        tables.append(dc.last_errors.clone())
    for t in tables[1:]:
        assert torch.equal(t, tables[0])
    assert torch.isfinite(tables[0]).all()
