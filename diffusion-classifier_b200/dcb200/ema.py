"""EMA -- the subset of ema_pytorch.EMA (ema-pytorch 0.7.7, requirements.txt:11) the reference relies on
(diffusion/diffusion_classifier.py:10,51-56,453,700): ``EMA(model, beta=, update_after_step=, update_every=)``,
``.ema_model`` / ``.online_model``, ``.update()`` and ``forward`` -> ``ema_model``.  State-dict keys match
(``online_model.*``, ``ema_model.*``, ``initted``, ``step``) so accelerate checkpoints of the reference load.

``update()`` follows ema_pytorch 0.7.7's schedule (restated -- the package is not installable offline, so this is
checked against the published formula only, tests/test_cpu_oracle.py::test_ema_update_schedule):

    step = self.step; self.step += 1
    if step % update_every:                 return
    if step <= update_after_step:           ema <- online (copy; ``initted`` untouched);  return
    if not initted:                         ema <- online;  initted = True
    decay = get_current_decay()             # warm-up: 1 - (1 + epoch / inv_gamma) ** -power, clamped to [min_value, beta],
    ema <- lerp(ema, online, 1 - decay)     #          epoch = max(self.step - update_after_step - 1, 0); 0 when epoch <= 0

for every floating-point parameter AND buffer (integer buffers are not averaged)."""
import copy

import torch
import torch.nn as nn


class EMA(nn.Module):
    def __init__(self, model, beta=0.9999, update_after_step=100, update_every=10, inv_gamma=1.0, power=2 / 3,
                 min_value=0.0, **unused):
        super().__init__()
        self.beta, self.update_after_step, self.update_every = beta, update_after_step, update_every
        self.inv_gamma, self.power, self.min_value = inv_gamma, power, min_value
        self.online_model = model
        self.ema_model = copy.deepcopy(model)
        self.ema_model.requires_grad_(False)
        self.register_buffer("initted", torch.tensor(False))
        self.register_buffer("step", torch.tensor(0))

    @staticmethod
    def _float_tensors(module):
        for group in (module.named_parameters(), module.named_buffers()):
            for name, t in group:
                if t.dtype.is_floating_point or t.dtype.is_complex:
                    yield name, t

    def get_current_decay(self):
        epoch = max(int(self.step.item()) - self.update_after_step - 1, 0)
        if epoch <= 0:
            return 0.0
        value = 1.0 - (1.0 + epoch / self.inv_gamma) ** -self.power
        return min(max(value, self.min_value), self.beta)

    @torch.no_grad()
    def copy_params_from_model_to_ema(self):
        for (_, pe), (_, po) in zip(self._float_tensors(self.ema_model), self._float_tensors(self.online_model)):
            pe.copy_(po)

    @torch.no_grad()
    def update(self):
        step = int(self.step.item())
        self.step += 1
        if step % self.update_every != 0:
            return
        if step <= self.update_after_step:
            self.copy_params_from_model_to_ema()
            return
        if not bool(self.initted.item()):
            self.copy_params_from_model_to_ema()
            self.initted.fill_(True)
        decay = self.get_current_decay()
        for (_, pe), (_, po) in zip(self._float_tensors(self.ema_model), self._float_tensors(self.online_model)):
            pe.lerp_(po, 1.0 - decay)

    def forward(self, *args, **kwargs):
        return self.ema_model(*args, **kwargs)
