"""ORACLE (test infrastructure only).  Imports the reference's OWN Python sources verbatim.

/root/reference is a pure-Python tree whose hot-path loop (diffusion/diffusion_classifier.py) imports
cleanly once its absent third-party dependencies are stubbed in ``sys.modules``:

  comet_ml, accelerate      -> inert placeholders (logging / launcher; never reached by classify)
  ema_pytorch.EMA           -> deepcopy + forward passthrough (ema-pytorch 0.7.7 semantics that classify
                               relies on: ``self.ema(...)`` runs ``ema_model`` ; diffusion_classifier.py:51-56,700)
  diffusers                 -> oracle/diffusers_restated.py (UNet2DConditionModel / DiTTransformer2DModel)
  pywt                      -> oracle/haar.py

Nothing is copied: the reference modules are executed from where they lie.  /root/reference does not
exist on the GPU box, so only CPU tests here and oracle/make_golden.py (which writes tests/golden/) use
this loader; everything that must run on the GPU box uses oracle/loop.py + the committed fixtures.
"""
from __future__ import annotations

import copy
import importlib
import os
import sys
import types
from contextlib import contextmanager

import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get("DCB_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "diffusion", "diffusion_classifier.py"))


class _EMA(nn.Module):
    def __init__(self, model, beta=0.9999, update_after_step=100, update_every=10, **kw):
        super().__init__()
        self.online_model = model
        self.ema_model = copy.deepcopy(model)
        self.ema_model.requires_grad_(False)
        self.register_buffer("initted", torch.tensor(False))
        self.register_buffer("step", torch.tensor(0))

    def update(self):
        self.step += 1

    def forward(self, *a, **k):
        return self.ema_model(*a, **k)


def _install_stubs():
    from . import diffusers_restated, haar

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    placeholder = type("Placeholder", (), {"__init__": lambda self, *a, **k: None})
    if "comet_ml" not in sys.modules:
        mod("comet_ml", Experiment=placeholder, ExistingExperiment=placeholder)
    if "accelerate" not in sys.modules:
        mod("accelerate", Accelerator=placeholder)
    if "ema_pytorch" not in sys.modules:
        mod("ema_pytorch", EMA=_EMA)
    if "pywt" not in sys.modules:
        sys.modules["pywt"] = haar.as_pywt_module()
    if "diffusers" not in sys.modules:
        d = mod("diffusers", UNet2DConditionModel=diffusers_restated.UNet2DConditionModel,
                UNet2DModel=diffusers_restated.UNet2DModel,
                DiTTransformer2DModel=diffusers_restated.DiTTransformer2DModel)
        d.__dcb_oracle_shim__ = True


def load_reference():
    """Returns a namespace with the reference's verbatim classes / functions."""
    if not reference_available():
        raise FileNotFoundError(f"reference tree not found at {REFERENCE_ROOT}")
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # the reference uses top-level package names (diffusion, nets, utils); make sure none is shadowed
    for name in ("diffusion", "nets", "utils"):
        m = sys.modules.get(name)
        if m is not None and not (getattr(m, "__file__", None) or REFERENCE_ROOT).startswith(REFERENCE_ROOT):
            paths = list(getattr(m, "__path__", []))
            if not any(p.startswith(REFERENCE_ROOT) for p in paths):
                raise ImportError(f"module {name!r} already imported from elsewhere: {m}")
    dc = importlib.import_module("diffusion.diffusion_classifier")
    unet = importlib.import_module("nets.unet")
    dit = importlib.import_module("nets.dit")
    wav = importlib.import_module("utils.wavelet")
    return types.SimpleNamespace(
        DiffusionClassifier=dc.DiffusionClassifier, UNetCondition2D=unet.UNetCondition2D, DiT=dit.DiT,
        wavelet_dec_2=wav.wavelet_dec_2, wavelet_enc_2=wav.wavelet_enc_2, log=dc.log, module=dc)


class Config:
    """Duck-typed config like the experiments' TrainingConfig (experiments/cifar10/train.py:24-38):
    missing keys read as None.  Dunder lookups raise so copy.deepcopy works."""

    def __init__(self, **kw):
        self.__dict__["_d"] = dict(kw)

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return self.__dict__["_d"].get(name)

    def __setattr__(self, name, value):
        self.__dict__["_d"][name] = value

    def as_dict(self):
        return dict(self.__dict__["_d"])


@contextmanager
def injected_noise(dc_obj, t_all, eps_all):
    """Run the reference's classify with pre-drawn (t, eps): patches torch.rand (:688) and the instance's
    diffuse (:100-117) and snoops torch.topk (:720) to recover the per-stage mean error tables."""
    state = {"j": 0, "stage_means": []}
    real_rand, real_topk = torch.rand, torch.topk

    def fake_rand(*size, **kw):
        n = size[0] if len(size) == 1 and isinstance(size[0], int) else None
        if n is not None and n == t_all.shape[1] and state["j"] < t_all.shape[0]:
            return t_all[state["j"]].clone().cpu()
        return real_rand(*size, **kw)

    def fake_diffuse(x, alpha_t, sigma_t):
        eps = eps_all[state["j"]].to(x.device)
        state["j"] += 1
        return alpha_t * x + sigma_t * eps, eps

    def snoop_topk(inp, *a, **k):
        state["stage_means"].append(inp.detach().clone())
        return real_topk(inp, *a, **k)

    dc_obj.diffuse = fake_diffuse
    torch.rand, torch.topk = fake_rand, snoop_topk
    try:
        yield state
    finally:
        torch.rand, torch.topk = real_rand, real_topk
        del dc_obj.diffuse
