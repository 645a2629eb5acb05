"""CPU tests of the SURVEY 8(f) rows: the restated sampler / loss oracle against the reference (golden fixtures written by
its verbatim code, and a direct A/B where /root/reference exists), the metrics module against the reference's verbatim
utils/metrics.py, and accelerate-layout checkpoints (host logic only: no kernels run here)."""
import os

import numpy as np
import pytest
import torch

from helpers import TINY_DIT, TINY_UNET, base_cfg
from test_cpu_oracle import _build, checksum, gold

SL_FIX = {
    "sample_loss_unet_tiny.npz": ("unet", TINY_UNET, dict(pred_param="eps", schedule="cosine", noise_d=16, image_size=16,
                                                           encoder_type="nn", classes=4, sampling_steps=4, cfg_w=1.5)),
    "sample_loss_dit_tiny.npz": ("dit", TINY_DIT, dict(pred_param="v", schedule="shifted_cosine", noise_d=16,
                                                        image_size=32, encoder_type="DiT", classes=3, sampling_steps=3,
                                                        cfg_w=0.5)),
}


@pytest.mark.parametrize("name", list(SL_FIX))
def test_sample_and_loss_oracle_reproduce_reference_golden(name):
    from oracle import loop
    kind, arch, kw = SL_FIX[name]
    g = gold(name)
    cfg = base_cfg(**kw)
    net, enc = _build(kind, arch, cfg, int(g["seed"]), float(g["factor"]))
    if abs(checksum(net) - float(g["checksum"])) > 1e-6 * float(g["checksum"]):
        pytest.skip("torch default-init stream differs from the build that wrote the fixture")
    x, text = torch.from_numpy(g["x"]), torch.from_numpy(g["text"])
    den = lambda z, lam, encoder_hidden_states: net(z, lam, encoder_hidden_states)[0]  # noqa: E731
    out = loop.sample_oracle(den, enc, cfg, x, text, from_t=float(g["from_t"]), z_init=torch.from_numpy(g["z_init"]),
                             noise_all=torch.from_numpy(g["noise_all"]))
    assert np.allclose(out.numpy(), g["sample"], atol=1e-5)
    with torch.no_grad():
        l = loop.loss_oracle(lambda x, noise_labels, encoder_hidden_states: net(x, noise_labels, encoder_hidden_states)[0],
                             enc, cfg, x, text, t=torch.from_numpy(g["t"]), eps=torch.from_numpy(g["eps"]))
    assert abs(float(l) - float(g["loss"])) < 1e-5 * float(g["loss"])


def test_sample_and_loss_oracle_equal_verbatim_reference_with_toy_backbone():
    """direct A/B: both sides draw from the default CPU generator in the same order, so equal seeds give equal noise."""
    from oracle import loop
    from oracle.reference_loader import Config, load_reference, reference_available
    if not reference_available():
        pytest.skip("/root/reference not present (GPU box)")
    ref = load_reference()

    class Toy(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.randn(8, 3))
            self.config = type("c", (), {"encoder_hid_dim": 8})()

        def forward(self, x, noise_labels, encoder_hidden_states):
            s = torch.tanh(encoder_hidden_states[:, 0] @ self.w)
            return x * s[:, :, None, None] + torch.tanh(noise_labels).view(-1, 1, 1, 1) * 0.1

    for pred, sched, w, from_t in (("eps", "cosine", 0.0, 1), ("v", "shifted_cosine", 2.0, 1), ("eps", "cosine", 1.0, 0.6)):
        torch.manual_seed(0)
        cfg = Config(pred_param=pred, schedule=sched, noise_d=16, image_size=8, cfg_w=w, ema_beta=0.99, ema_warmup=0,
                     ema_update_freq=1, encoder_type="nn", classes=5, sampling_steps=6)
        dc = ref.DiffusionClassifier(Toy(), cfg).eval()
        x = torch.rand(4, 3, 8, 8) * 2 - 1
        text = torch.randint(0, 5, (4,))
        torch.manual_seed(77)
        y_ref = dc.sample(x, text, from_t=from_t)
        torch.manual_seed(77)
        y = loop.sample_oracle(lambda z, lam, encoder_hidden_states: dc.ema(z, lam, encoder_hidden_states=encoder_hidden_states),
                               dc.encoder, cfg, x, text, from_t=from_t)
        assert torch.equal(y, y_ref)
        torch.manual_seed(78)
        l_ref = dc.loss(x, text)
        torch.manual_seed(78)
        l = loop.loss_oracle(dc.model, dc.encoder, cfg, x, text)
        assert torch.equal(l, l_ref)
        # ddpm_sampler_step in isolation
        z, p, u = torch.randn(4, 3, 8, 8), torch.randn(4, 3, 8, 8), torch.randn(4, 3, 8, 8)
        lt, ls = torch.tensor([-1.3]), torch.tensor([0.4])
        mu_ref, var_ref = dc.ddpm_sampler_step(z, p, u, lt, ls)
        mu, var = loop.ddpm_sampler_step_oracle(cfg, z, p, u, lt, ls)
        assert torch.equal(mu, mu_ref) and torch.equal(var, var_ref)


def test_metrics_match_reference_semantics():
    """dcb200.metrics vs the reference's verbatim utils/metrics.py on the same (prediction, batch) stream."""
    from dcb200 import metrics as M
    g = torch.Generator().manual_seed(0)
    stream = [(torch.randint(0, 2, (7,), generator=g), {"prompt": torch.randint(0, 2, (7,), generator=g)})
              for _ in range(5)]
    mine = [M.Accuracy("accuracy"), M.Precision("precision"), M.Recall("recall"), M.F1("f1")]
    for m in mine:
        m.set_device(torch.device("cpu"))
        for out in stream:
            m.update(out)
    yp = torch.cat([s[0] for s in stream])
    yt = torch.cat([s[1]["prompt"] for s in stream])
    tp, fp, fn = ((yp == 1) & (yt == 1)).sum(), ((yp == 1) & (yt == 0)).sum(), ((yp == 0) & (yt == 1)).sum()
    want = {"accuracy": (yp == yt).float().mean(), "precision": tp / (tp + fp), "recall": tp / (tp + fn),
            "f1": 2 * tp / (2 * tp + fp + fn)}
    for m in mine:
        assert abs(float(m.get_output()[m.name]) - float(want[m.name])) < 1e-6
    from oracle.reference_loader import reference_available
    if reference_available():
        import importlib.util
        spec = importlib.util.spec_from_file_location("_ref_metrics", "/root/reference/utils/metrics.py")
        R = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(R)
        for mine_m, ref_m in zip(mine, [R.Accuracy("accuracy"), R.Precision("precision"), R.Recall("recall"), R.F1("f1")]):
            for out in stream:
                ref_m.update(out)
            assert abs(float(ref_m.get_output()[ref_m.name]) - float(mine_m.get_output()[mine_m.name])) < 1e-6
    empty = M.Precision("precision")
    assert empty.get_output()["precision"] == 0.0       # reference: 0.0 when the denominator is 0
    mine[0].reset()
    assert int(mine[0].total) == 0


def test_checkpoint_accelerate_layout_roundtrip(tmp_path):
    """save_checkpoint writes model.safetensors / model_1.safetensors (EMA) / model_2.safetensors (encoder) +
    experiment_state.pth (what accelerate.save_state + diffusion_classifier.py:727-766 write); a checkpoint produced from
    the ORACLE modules (diffusers key schema, ema_pytorch buffers) loads into the product classes."""
    import dcb200
    from oracle import diffusers_restated as dr
    from oracle.reference_loader import _EMA
    from safetensors.torch import save_file
    cfg = base_cfg(classes=3)
    torch.manual_seed(0)
    a = dcb200.DiffusionClassifier(dcb200.UNetCondition2D(**TINY_UNET), cfg)
    d = str(tmp_path / "ck")
    a.save_checkpoint(d, epoch=7, best_metric=0.5, experiment_key="abc")
    assert sorted(os.listdir(d)) == ["experiment_state.pth", "model.safetensors", "model_1.safetensors",
                                     "model_2.safetensors"]
    torch.manual_seed(1)
    b = dcb200.DiffusionClassifier(dcb200.UNetCondition2D(**TINY_UNET), cfg)
    assert b.load_checkpoint(d) == (7, 0.5, "abc")
    for (k, v), (k2, v2) in zip(a.state_dict().items(), b.state_dict().items()):
        assert k == k2 and torch.equal(v, v2)
    # reference-side writer: oracle U-Net + ema_pytorch-shaped EMA + nn.Embedding, saved as accelerate would
    torch.manual_seed(2)
    o = dr.UNet2DConditionModel(**TINY_UNET)
    ema = _EMA(o)
    with torch.no_grad():
        for p in ema.ema_model.parameters():
            p.add_(0.01)
    enc = torch.nn.Embedding(4, TINY_UNET["encoder_hid_dim"])
    d2 = str(tmp_path / "ref")
    os.makedirs(d2)
    for i, m in enumerate((o, ema, enc)):
        save_file({k: v.contiguous().clone() for k, v in m.state_dict().items()},
                  os.path.join(d2, "model.safetensors" if i == 0 else f"model_{i}.safetensors"))
    assert b.load_checkpoint(d2) == (0, None, None)
    assert torch.equal(b.model.conv_in.weight, o.conv_in.weight)
    assert torch.equal(b.ema.ema_model.conv_in.weight, ema.ema_model.conv_in.weight)
    assert not torch.equal(b.ema.ema_model.conv_in.weight, b.model.conv_in.weight)
    assert torch.equal(b.encoder.weight, enc.weight)
    with pytest.raises(FileNotFoundError):
        b.load_checkpoint(str(tmp_path / "missing"))


def test_sampler_coefficients_match_step_oracle():
    """host-side coefficient table fed to dcb_ddpm_step == what ddpm_sampler_step (:189-205) computes per evaluation,
    including the reference's extra "final step" row."""
    import dcb200
    from oracle import loop
    cfg = base_cfg(classes=3, sampling_steps=5, cfg_w=1.25, schedule="shifted_cosine", noise_d=16, image_size=32)
    dc = dcb200.DiffusionClassifier(dcb200.UNetCondition2D(**TINY_UNET), cfg)
    coef, lt = dc._sampler_coefs(0.8, torch.device("cpu"))
    assert coef.shape == (6, 8) and lt.shape == (6,)
    sched = loop.schedule_fn(cfg)
    steps = torch.linspace(0.8, 0.0, 6)
    pairs = [(steps[i], steps[i + 1]) for i in range(5)] + [(steps[-2], steps[-1])]
    for i, (ut, us) in enumerate(pairs):
        l_t, l_s = sched(ut).unsqueeze(0), sched(us).unsqueeze(0)
        assert torch.allclose(lt[i], l_t[0])
        z = torch.zeros(1, 1, 1, 1)
        _, var = loop.ddpm_sampler_step_oracle(cfg, z, z, z, l_t, l_s)
        assert torch.allclose(coef[i, 5] ** 2, var[0], rtol=1e-5, atol=1e-12)
        assert torch.allclose(coef[i, 0], -torch.special.expm1(l_t - l_s)[0])
        assert float(coef[i, 6]) == 1.25


def _metrics_gloo_worker(rank, world, port, out):
    import torch.distributed as dist
    from dcb200 import metrics as M
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(100 + rank)
    mets = [M.Accuracy("accuracy"), M.F1("f1")]
    for i in range(3):      # replicas: every rank scores its own batches (inference() without dcb_shard)
        yp, yt = torch.randint(0, 2, (5,), generator=g), torch.randint(0, 2, (5,), generator=g)
        for m in mets:
            m.update((yp, {"prompt": yt}))
    for m in mets:
        m.sync_across_processes(None)          # one all-reduce per metric
    if rank == 0:
        torch.save({m.name: float(m.get_output()[m.name]) for m in mets}, out)
    dist.destroy_process_group()


def test_metrics_sync_across_processes_gloo_world2(tmp_path):
    """replica mode of inference(): per-rank counters are summed by one all-reduce (utils/metrics.py's accelerator.reduce)."""
    import torch.multiprocessing as mp
    out = str(tmp_path / "m.pt")
    port = 31500 + os.getpid() % 2000
    mp.spawn(_metrics_gloo_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    yp, yt = [], []
    for rank in range(2):
        g = torch.Generator().manual_seed(100 + rank)
        for i in range(3):
            yp.append(torch.randint(0, 2, (5,), generator=g))
            yt.append(torch.randint(0, 2, (5,), generator=g))
    yp, yt = torch.cat(yp), torch.cat(yt)
    tp, fp, fn = ((yp == 1) & (yt == 1)).sum(), ((yp == 1) & (yt == 0)).sum(), ((yp == 0) & (yt == 1)).sum()
    assert abs(got["accuracy"] - float((yp == yt).float().mean())) < 1e-6
    assert abs(got["f1"] - float(2 * tp / (2 * tp + fp + fn))) < 1e-6


def test_fold_upsample_weights_identity_on_cpu():
    """host algebra behind dcb_gemm_desc.up_phase: nearest-2x upsample + conv3x3(pad 1) == four 2x2-tap convs over the
    low-resolution input whose outputs interleave as (2y + a, 2x + b) -- checked with torch on the CPU in float64."""
    import torch.nn.functional as F
    from dcb200 import engine as E
    torch.manual_seed(0)
    N, C, Co, H, W = 2, 5, 7, 6, 9
    x = torch.randn(N, C, H, W, dtype=torch.float64)
    w = torch.randn(Co, C, 3, 3, dtype=torch.float64)
    b = torch.randn(Co, dtype=torch.float64)
    ref = F.conv2d(F.interpolate(x, scale_factor=2.0, mode="nearest"), w, b, padding=1)
    ph = E.fold_upsample_weights(w)
    assert len(ph) == 4 and ph[0].shape == (Co, 4 * C)
    out = torch.empty_like(ref)
    xp = F.pad(x, (1, 1, 1, 1))
    for a in range(2):
        for bb in range(2):
            wk = ph[2 * a + bb].reshape(Co, 2, 2, C).permute(0, 3, 1, 2)      # [(ty, tx, c)] -> OIHW
            out[:, :, a::2, bb::2] = F.conv2d(xp[:, :, a:a + H + 1, bb:bb + W + 1], wk, b)
    assert torch.allclose(out, ref, atol=1e-12)
    # the segment offsets engine.upsample_conv passes: phase (a, b), tap (ty, tx) reads source pixel (y - 1 + a + ty, x - 1 + b + tx)
    y0, x0, a, bb = 3, 4, 1, 0
    acc = b.clone()
    for ty in range(2):
        for tx in range(2):
            yy, xx = y0 - 1 + a + ty, x0 - 1 + bb + tx
            if 0 <= yy < H and 0 <= xx < W:
                acc += ph[2 * a + bb].reshape(Co, 2, 2, C)[:, ty, tx] @ x[0, :, yy, xx]
    assert torch.allclose(acc, ref[0, :, 2 * y0 + a, 2 * x0 + bb], atol=1e-12)
