"""EMA -- the subset of ema_pytorch.EMA (ema-pytorch 0.7.7) the reference relies on
(diffusion/diffusion_classifier.py:10,51-56,453,700): ``EMA(model, beta=, update_after_step=, update_every=)``,
``.ema_model`` / ``.online_model``, ``.update()`` and ``forward`` -> ``ema_model``.  State-dict keys match
(``online_model.*``, ``ema_model.*``, ``initted``, ``step``) so accelerate checkpoints of the reference load."""
import copy

import torch
import torch.nn as nn


class EMA(nn.Module):
    def __init__(self, model, beta=0.9999, update_after_step=100, update_every=10, **unused):
        super().__init__()
        self.beta, self.update_after_step, self.update_every = beta, update_after_step, update_every
        self.online_model = model
        self.ema_model = copy.deepcopy(model)
        self.ema_model.requires_grad_(False)
        self.register_buffer("initted", torch.tensor(False))
        self.register_buffer("step", torch.tensor(0))

    @torch.no_grad()
    def update(self):
        step = int(self.step.item())
        self.step += 1
        if step % self.update_every != 0:
            return
        if step <= self.update_after_step or not bool(self.initted.item()):
            for pe, po in zip(self.ema_model.parameters(), self.online_model.parameters()):
                pe.copy_(po)
            self.initted.fill_(True)
            return
        for pe, po in zip(self.ema_model.parameters(), self.online_model.parameters()):
            pe.lerp_(po, 1.0 - self.beta)

    def forward(self, *args, **kwargs):
        return self.ema_model(*args, **kwargs)
