"""ORACLE (test infrastructure only).  From-scratch restatement of the reference's ELBO classification
loop, for use where /root/reference is absent (the GPU box).  Checked against the reference's VERBATIM
code in tests/test_oracle_loop.py (here, where /root/reference exists) and against tests/golden/*.npz.

Follows /root/reference/diffusion/diffusion_classifier.py:
  log                       :14-15      logsnr_schedule_cosine(_shifted)   :119-161
  diffuse                   :100-117    encode_text_prompt                 :83-98
  classify                  :657-725
Noise injection: the reference draws t with the CPU generator (torch.rand(BS), :688) and eps with
randn_like on x.device (:113).  ``classify_oracle`` draws them the same way unless ``t_all`` / ``eps_all``
([T,BS] / [T,BS,C,H,W]) are supplied, in which case those are used (identical pre-drawn noise for parity).
"""
import math

import torch


def log_clamped(t, eps=1e-20):
    return torch.log(t.clamp(min=eps))


def logsnr_schedule_cosine(t, noise_d, image_d, logsnr_min=-15, logsnr_max=15):
    logsnr_max = logsnr_max + math.log(noise_d / image_d)
    logsnr_min = logsnr_min + math.log(noise_d / image_d)
    t_min = math.atan(math.exp(-0.5 * logsnr_max))
    t_max = math.atan(math.exp(-0.5 * logsnr_min))
    return -2 * log_clamped(torch.tan(t_min + t * (t_max - t_min)))


def logsnr_schedule_cosine_shifted(t, noise_d, image_d):
    return logsnr_schedule_cosine(t, noise_d, image_d) + 2 * math.log(noise_d / image_d)


def schedule_fn(config):
    if config.schedule == "cosine":
        return lambda t: logsnr_schedule_cosine(t, config.noise_d, config.image_size)
    assert config.schedule == "shifted_cosine"
    return lambda t: logsnr_schedule_cosine_shifted(t, config.noise_d, config.image_size)


@torch.no_grad()
def classify_oracle(denoiser, encoder, config, x, text=None, fast=False, t_all=None, eps_all=None,
                    return_errors=False):
    """denoiser(x=, noise_labels=, encoder_hidden_states=) -> prediction; encoder = nn.Embedding or None (DiT)."""
    assert len(config.evaluation_per_stage) == config.n_stages
    assert len(config.n_keep_per_stage) == config.n_stages
    assert config.n_keep_per_stage[-1] == 1
    assert 2 <= config.n_fast_classes <= config.classes
    sched = schedule_fn(config)
    per_stage = [0] + list(config.evaluation_per_stage)
    BS = x.shape[0]
    errors = torch.full((BS, config.classes, per_stage[-1]), torch.inf).to(x.device)
    if fast:
        text = text.view(-1, 1)
        classes = torch.arange(config.classes).repeat(BS, 1).to(x.device)
        wrong = classes[(classes == text) == False].view(BS, -1)  # noqa: E712
        sel = torch.randint(0, wrong.shape[1], (BS, config.n_fast_classes - 1)).to(x.device)
        classes = torch.cat((text, torch.gather(wrong, 1, sel)), dim=1)
    else:
        classes = torch.arange(config.classes).repeat(BS, 1).to(x.device)
    for i in range(config.n_stages):
        for j in range(per_stage[i], per_stage[i + 1]):
            t = torch.rand(BS) if t_all is None else t_all[j].cpu()
            logsnr = sched(t).to(x.device)
            alpha = torch.sqrt(torch.sigmoid(logsnr)).view(-1, 1, 1, 1)
            sigma = torch.sqrt(torch.sigmoid(-logsnr)).view(-1, 1, 1, 1)
            eps = torch.randn_like(x) if eps_all is None else eps_all[j].to(x.device)
            z = alpha * x + sigma * eps
            for c in range(classes.shape[1]):
                lab = classes[:, c]
                emb = encoder(lab).unsqueeze(1) if encoder is not None else lab
                pred = denoiser(x=z, noise_labels=logsnr, encoder_hidden_states=emb)
                eps_pred = sigma * z + alpha * pred if config.pred_param == "v" else pred
                err = torch.norm((eps_pred - eps).view(BS, -1), dim=1, p=2) ** 2
                errors[torch.arange(BS, device=x.device), lab, j] = err
        stage_err = errors[:, :, :per_stage[i + 1]].mean(dim=2)
        _, classes = torch.topk(stage_err, config.n_keep_per_stage[i], dim=1, largest=False)
    assert classes.shape[1] == 1
    return (classes[:, 0], errors) if return_errors else classes[:, 0]


# ---- next rows f2 / f4: DDPM sampler with classifier-free guidance, and the training loss forward -------
def ddpm_sampler_step_oracle(config, z_t, pred, u_pred, logsnr_t, logsnr_s):
    """diffusion_classifier.py:176-207 restated: returns (mu, variance)."""
    c = -torch.special.expm1(logsnr_t - logsnr_s)
    alpha_t, alpha_s = torch.sqrt(torch.sigmoid(logsnr_t)), torch.sqrt(torch.sigmoid(logsnr_s))
    sigma_t, sigma_s = torch.sqrt(torch.sigmoid(-logsnr_t)), torch.sqrt(torch.sigmoid(-logsnr_s))
    w = config.cfg_w
    pred = (1 + w) * pred - w * u_pred
    x_pred = alpha_t * z_t - sigma_t * pred if config.pred_param == "v" else (z_t - sigma_t * pred) / alpha_t
    x_pred = torch.clamp(x_pred, -1, 1)
    mu = alpha_s * (z_t * (1 - c) / alpha_t + c * x_pred)
    return mu, (sigma_s ** 2) * c


@torch.no_grad()
def sample_oracle(denoiser, encoder, config, x, text, from_t=1, z_init=None, noise_all=None):
    """diffusion_classifier.py:209-293 restated.  denoiser(z, logsnr[1], encoder_hidden_states=) -> prediction;
    encoder = nn.Embedding (null token id = config.classes) or None (DiT: labels pass through).
    z_init [B,C,H,W]: the initial state (the reference draws randn, or diffuses x to from_t);
    noise_all [sampling_steps,B,C,H,W]: the per-step randn_like draws.  Note the reference evaluates the denoiser
    sampling_steps + 1 times: the loop's last iteration already steps to t = 0, then the "final step" re-evaluates at
    steps[-2] from that state and returns clip(mu)."""
    sched = schedule_fn(config)
    if z_init is not None:
        z_t = z_init.clone()
    elif from_t == 1:
        z_t = torch.randn(x.shape).to(x.device)
    else:
        logsnr = sched(torch.ones(x.shape[0]) * from_t).to(x.device)
        a = torch.sqrt(torch.sigmoid(logsnr)).view(-1, 1, 1, 1)
        s = torch.sqrt(torch.sigmoid(-logsnr)).view(-1, 1, 1, 1)
        z_t = a * x + s * torch.randn_like(x)
    null = torch.full_like(text, config.classes)
    emb = encoder(text).unsqueeze(1) if encoder is not None else text
    nemb = encoder(null).unsqueeze(1) if encoder is not None else null
    steps = torch.linspace(from_t, 0.0, config.sampling_steps + 1)
    for i in range(len(steps) - 1):
        lt, ls = sched(steps[i]).to(x.device).unsqueeze(0), sched(steps[i + 1]).to(x.device).unsqueeze(0)
        pred = denoiser(z_t, lt, encoder_hidden_states=emb)
        u_pred = denoiser(z_t, lt, encoder_hidden_states=nemb)
        mu, var = ddpm_sampler_step_oracle(config, z_t, pred, u_pred, lt, ls)
        n = torch.randn_like(mu) if noise_all is None else noise_all[i].to(x.device)
        z_t = mu + n * torch.sqrt(var)
    l1, l0 = sched(steps[-2]).to(x.device).unsqueeze(0), sched(steps[-1]).to(x.device).unsqueeze(0)
    pred = denoiser(z_t, l1, encoder_hidden_states=emb)
    u_pred = denoiser(z_t, l1, encoder_hidden_states=nemb)
    x_pred, _ = ddpm_sampler_step_oracle(config, z_t, pred, u_pred, l1, l0)
    return torch.clamp(x_pred, -1, 1)


def loss_oracle(denoiser, encoder, config, x, text, t=None, eps=None):
    """diffusion_classifier.py:295-344 restated (min-SNR weighted eps-MSE): denoiser(x=, noise_labels=,
    encoder_hidden_states=).  t [B] / eps [B,C,H,W] inject the reference's torch.rand / randn_like draws."""
    sched = schedule_fn(config)
    t = torch.rand(x.shape[0]) if t is None else t.cpu()
    emb = (encoder(text).unsqueeze(1) if encoder is not None else text) if text is not None else None
    logsnr = sched(t).to(x.device)
    alpha = torch.sqrt(torch.sigmoid(logsnr)).view(-1, 1, 1, 1)
    sigma = torch.sqrt(torch.sigmoid(-logsnr)).view(-1, 1, 1, 1)
    eps = torch.randn_like(x) if eps is None else eps
    z = alpha * x + sigma * eps
    pred = denoiser(x=z, noise_labels=logsnr, encoder_hidden_states=emb)
    eps_pred = sigma * z + alpha * pred if config.pred_param == "v" else pred
    snr = torch.exp(logsnr).clamp_(max=5)
    weight = 1 / (1 + snr) if config.pred_param == "v" else 1 / snr
    return torch.mean(weight.view(-1, 1, 1, 1) * (eps_pred - eps) ** 2)
