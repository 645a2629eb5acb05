"""ORACLE (test infrastructure only).  Writes tests/golden/*.npz by running the reference's OWN code from
/root/reference (verbatim classify loop / schedule / wrappers / wavelet module, see reference_loader.py) around
the restated diffusers denoisers.  Run here (where /root/reference exists):  python -m oracle.make_golden

The fixtures hold inputs' seeds, pre-drawn noise and the reference outputs; weights are NOT stored -- they are
re-created from ``torch.manual_seed(seed)`` default init (same torch build on the GPU box) and guarded by a checksum.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import SMALL_UNET, TINY_DIT, TINY_UNET  # noqa: E402
from oracle.reference_loader import Config, injected_noise, load_reference  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def checksum(module):
    return float(sum(p.detach().double().abs().sum() for p in module.parameters()))


def amplify_class_signal(dc, kind, factor):
    """random-init nets barely react to the class; scale the class pathway so margins are meaningful."""
    with torch.no_grad():
        if kind == "unet":
            dc.encoder.weight.mul_(factor)
        else:
            for m in (dc.model, dc.ema.ema_model):
                for b in m.transformer_blocks:
                    b.norm1.emb.class_embedder.embedding_table.weight.mul_(factor)


def classify_fixture(ref, name, kind, arch, cfg_kw, BS, seed, factor):
    torch.manual_seed(seed)
    net = (ref.UNetCondition2D if kind == "unet" else ref.DiT)(**arch)
    cfg = Config(**cfg_kw)
    dc = ref.DiffusionClassifier(net, cfg).eval()
    amplify_class_signal(dc, kind, factor)
    T = cfg.evaluation_per_stage[-1]
    C, S = arch["in_channels"], arch["sample_size"]
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.rand(BS, C, S, S, generator=g) * 2 - 1
    t_all = torch.rand(T, BS, generator=g)
    eps_all = torch.randn(T, BS, C, S, S, generator=g)
    with injected_noise(dc, t_all, eps_all) as st:
        labels = dc.classify(x)
    np.savez_compressed(
        os.path.join(OUT, name), x=x.numpy(), t_all=t_all.numpy(), eps_all=eps_all.numpy(), labels=labels.numpy(),
        stage_means=np.stack([m.numpy() for m in st["stage_means"]]), seed=seed, factor=factor,
        checksum=checksum(dc.ema.ema_model), enc_checksum=checksum(dc.encoder) if dc.encoder is not None else 0.0)
    print(name, "labels", labels.tolist(), "final means", st["stage_means"][-1][0].tolist())


def sample_loss_fixture(ref, name, kind, arch, cfg_kw, BS, seed, factor, from_t=1):
    """verbatim DiffusionClassifier.sample (:209-293) and .loss (:295-344); the draws the reference makes with the
    default CPU generator (randn for z_T, randn_like per step; rand + randn_like for the loss) are replayed from the
    same seed and stored, so the product can be fed identical noise."""
    from oracle import loop
    torch.manual_seed(seed)
    net = (ref.UNetCondition2D if kind == "unet" else ref.DiT)(**arch)
    cfg = Config(**cfg_kw)
    dc = ref.DiffusionClassifier(net, cfg).eval()
    amplify_class_signal(dc, kind, factor)
    C, S = arch["in_channels"], arch["sample_size"]
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.rand(BS, C, S, S, generator=g) * 2 - 1
    text = torch.randint(0, cfg.classes, (BS,), generator=g)
    torch.manual_seed(seed + 2)
    out = dc.sample(x, text, from_t=from_t)
    torch.manual_seed(seed + 2)
    if from_t == 1:
        z_init = torch.randn(x.shape)
    else:
        lam = dc.schedule(torch.ones(BS) * from_t)
        a, sg = torch.sqrt(torch.sigmoid(lam)).view(-1, 1, 1, 1), torch.sqrt(torch.sigmoid(-lam)).view(-1, 1, 1, 1)
        z_init = a * x + sg * torch.randn_like(x)
    noise_all = torch.stack([torch.randn(x.shape) for _ in range(cfg.sampling_steps)])

    class Den(torch.nn.Module):
        def forward(self, x, noise_labels, encoder_hidden_states):
            return dc.ema(x, noise_labels, encoder_hidden_states=encoder_hidden_states)

    replay = loop.sample_oracle(Den(), dc.encoder, cfg, x, text, from_t=from_t, z_init=z_init, noise_all=noise_all)
    assert torch.equal(replay, out), "replayed draws must reproduce the reference's sample bit for bit"
    torch.manual_seed(seed + 3)
    loss = dc.loss(x, text)
    torch.manual_seed(seed + 3)
    t = torch.rand(BS)
    eps = torch.randn(x.shape)
    with torch.no_grad():
        l2 = loop.loss_oracle(lambda x, noise_labels, encoder_hidden_states: dc.model(
            x=x, noise_labels=noise_labels, encoder_hidden_states=encoder_hidden_states), dc.encoder, cfg, x, text, t=t,
            eps=eps)
    assert torch.equal(l2, loss.detach()), "replayed draws must reproduce the reference's loss bit for bit"
    np.savez_compressed(os.path.join(OUT, name), x=x.numpy(), text=text.numpy(), z_init=z_init.numpy(),
                        noise_all=noise_all.numpy(), sample=out.numpy(), t=t.numpy(), eps=eps.numpy(),
                        loss=float(loss), seed=seed, factor=factor, from_t=from_t, checksum=checksum(dc.ema.ema_model))
    print(name, "sample mean/std", float(out.mean()), float(out.std()), "loss", float(loss))


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = load_reference()
    base = dict(cfg_w=0.0, ema_beta=0.999, ema_warmup=0, ema_update_freq=1, n_fast_classes=2)

    # (1) schedule KATs straight from the reference's methods (diffusion_classifier.py:119-161)
    t = torch.tensor([0.0, 0.1, 0.25, 0.5, 0.75, 0.9, 1.0])
    rows = {}
    for tag, sched, nd, im in (("cos_32_32", "cosine", 32, 32), ("cos_64_256", "cosine", 64, 256),
                               ("shift_64_256", "shifted_cosine", 64, 256), ("shift_32_128", "shifted_cosine", 32, 128)):
        stub = torch.nn.Linear(1, 1)
        stub.config = type("c", (), {"encoder_hid_dim": 4})()
        dc = ref.DiffusionClassifier(stub, Config(pred_param="eps", schedule=sched, noise_d=nd, image_size=im,
                                                  encoder_type="nn", classes=2, **base))
        rows[tag] = dc.schedule(t).numpy()
    np.savez_compressed(os.path.join(OUT, "schedule_kat.npz"), t=t.numpy(), **rows)

    # (2) verbatim classify around the restated denoisers (2-stage pruning, eps and v, cosine and shifted)
    classify_fixture(ref, "classify_unet_tiny.npz", "unet", TINY_UNET,
                     dict(pred_param="eps", schedule="cosine", noise_d=16, image_size=16, encoder_type="nn", classes=4,
                          n_stages=2, evaluation_per_stage=[2, 4], n_keep_per_stage=[2, 1], **base), BS=3, seed=0,
                     factor=40.0)
    classify_fixture(ref, "classify_unet_small_v.npz", "unet", SMALL_UNET,
                     dict(pred_param="v", schedule="shifted_cosine", noise_d=16, image_size=32, encoder_type="nn",
                          classes=2, n_stages=1, evaluation_per_stage=[3], n_keep_per_stage=[1], **base), BS=2, seed=3,
                     factor=40.0)
    classify_fixture(ref, "classify_dit_tiny.npz", "dit", TINY_DIT,
                     dict(pred_param="v", schedule="shifted_cosine", noise_d=16, image_size=32, encoder_type="DiT",
                          classes=3, n_stages=2, evaluation_per_stage=[2, 3], n_keep_per_stage=[2, 1], **base), BS=2,
                     seed=5, factor=20.0)

    # (2b) verbatim sample (DDPM + classifier-free guidance) and loss
    sample_loss_fixture(ref, "sample_loss_unet_tiny.npz", "unet", TINY_UNET,
                        dict(pred_param="eps", schedule="cosine", noise_d=16, image_size=16, encoder_type="nn", classes=4,
                             sampling_steps=4, **{**base, "cfg_w": 1.5}), BS=2, seed=21, factor=40.0)
    sample_loss_fixture(ref, "sample_loss_dit_tiny.npz", "dit", TINY_DIT,
                        dict(pred_param="v", schedule="shifted_cosine", noise_d=16, image_size=32, encoder_type="DiT",
                             classes=3, sampling_steps=3, **{**base, "cfg_w": 0.5}), BS=2, seed=23, factor=20.0,
                        from_t=0.5)

    # (3) plain forwards through the reference's wrappers (nets/unet.py:186-195, nets/dit.py:49-51)
    torch.manual_seed(11)
    u = ref.UNetCondition2D(**SMALL_UNET).eval()
    g = torch.Generator().manual_seed(12)
    x = torch.randn(2, 12, 32, 32, generator=g)
    lam = torch.tensor([-3.0, 2.5])
    ehs = torch.randn(2, 1, 128, generator=g)
    with torch.no_grad():
        y = u(x, lam, encoder_hidden_states=ehs)
    np.savez_compressed(os.path.join(OUT, "unet_small_forward.npz"), x=x.numpy(), lam=lam.numpy(), ehs=ehs.numpy(),
                        y=y.numpy(), checksum=checksum(u))
    torch.manual_seed(13)
    d = ref.DiT(**TINY_DIT).eval()
    x = torch.randn(2, 3, 32, 32, generator=g)
    lab = torch.tensor([1, 0])
    with torch.no_grad():
        y = d(x, lam, lab)
    np.savez_compressed(os.path.join(OUT, "dit_tiny_forward.npz"), x=x.numpy(), lam=lam.numpy(), lab=lab.numpy(),
                        y=y.numpy(), checksum=checksum(d))

    # (4) the reference's wavelet module (utils/wavelet.py) over the restated pywt Haar
    img = torch.rand(3, 16, 24, generator=g) * 2 - 1
    w = ref.wavelet_dec_2(img)
    back = ref.wavelet_enc_2(w)
    np.savez_compressed(os.path.join(OUT, "haar_kat.npz"), img=img.numpy(), w=w.numpy(), back=back.numpy())
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
