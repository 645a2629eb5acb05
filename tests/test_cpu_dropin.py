"""The drop-in boundary (Python face): the reference's OWN files are executed, unmodified and from where they lie
(/root/reference -- nothing is copied), with ``diffusion-classifier_b200/dropin`` FIRST on sys.path, so that their
``from nets.unet import UNetCondition2D`` / ``from nets.dit import DiT`` / ``from diffusion.diffusion_classifier import
DiffusionClassifier`` / ``from utils.metrics import ...`` / ``from utils.wavelet import ...`` resolve to dcb200.

  * models/*.py (bare fragments with a free ``config``): every one of the six builds a dcb200 network whose kwargs equal
    dcb200.configs (what bench.py runs) and whose parameter counts are SURVEY App. D's;
  * experiments/chexpert-unet/inference.py and experiments/chexpert-dit/inference.py: ``main()`` runs top to bottom with only
    the dataset, accelerate and diffusers.optimization stubbed, up to the ``DiffusionClassifier.inference`` call (no GPU here;
    tests/test_gpu_h_dropin.py runs that call for real).

Each case runs in a fresh interpreter: the module names ``nets`` / ``diffusion`` / ``utils`` would collide with the
reference's verbatim packages that other CPU tests import through oracle/reference_loader.py.  Skipped where
/root/reference is absent (the GPU box)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("DCB_REFERENCE_ROOT", "/root/reference")
PKG = os.path.join(ROOT, "diffusion-classifier_b200")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "models")), reason="reference tree not present")

PRELUDE = f"""
import json, os, sys, types
sys.path[:0] = [{os.path.join(PKG, 'dropin')!r}, {PKG!r}]
import torch
import dcb200
def count(m):
    return sum(p.numel() for p in m.parameters())
"""


def _run(body, env=None):
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([sys.executable, "-c", PRELUDE + body], capture_output=True, text=True, timeout=600, env=e)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    return json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])


# fragment, the config keys it reads, dcb200.configs name (or None), parameter count (DiT: 8a6)
FRAGMENTS = [
    ("unet-128.py", dict(image_size=128, image_channels=3, wavelet_transform=False), "UNET128", 275819523),
    ("unet-256.py", dict(image_size=256, image_channels=3, wavelet_transform=False), "UNET256", 285.7),
    ("ipmsa-5-dwt-unet.py", dict(image_size=256, image_channels=10, wavelet_transform=True), "IPMSA5_DWT_UNET", 242062888),
    ("chexpert-256-dit-b4.py", dict(image_size=256, image_channels=3, wavelet_transform=False, patch_size=4), "DIT_B4_256",
     147476784),
    ("ipmsa-5-unet.py", dict(image_size=256, image_channels=10, wavelet_transform=False), None, 473.2),
    ("chexpert-256-unet-dwt-healthysick.py", dict(image_size=256, image_channels=3, wavelet_transform=True), None, None),
]


@pytest.mark.parametrize("frag,cfg,name,params", FRAGMENTS, ids=[f[0] for f in FRAGMENTS])
def test_reference_model_fragment_builds_the_dcb200_network(frag, cfg, name, params):
    body = f"""
from dcb200 import configs
config = configs.Config(**{cfg!r})
ns = {{"config": config}}
src = open({os.path.join(REF, 'models', frag)!r}).read()
with torch.device("meta"):
    exec(compile(src, {frag!r}, "exec"), ns)
net = ns.get("unet") or ns.get("dit")
kw = {{k: getattr(net.config, k) for k in vars(net.config)}}
out = dict(cls=type(net).__module__ + "." + type(net).__name__, params=count(net),
           sample_size=net.config.sample_size, in_channels=net.config.in_channels)
name = {name!r}
if name:
    with torch.device("meta"):
        mine = (dcb200.DiT if "dit" in {frag!r} else dcb200.UNetCondition2D)(**getattr(configs, name))
    out["same_config"] = all(getattr(mine.config, k) == v for k, v in kw.items())
    out["same_keys"] = sorted(mine.state_dict()) == sorted(net.state_dict())
    out["same_shapes"] = all(tuple(a.shape) == tuple(b.shape) for a, b in zip(mine.state_dict().values(),
                                                                               net.state_dict().values()))
print(json.dumps(out))
"""
    out = _run(body)
    assert out["cls"] in ("dcb200.unet.UNetCondition2D", "dcb200.dit.DiT")
    if name:
        assert out["same_config"] and out["same_keys"] and out["same_shapes"]
    if isinstance(params, int):
        assert out["params"] == params
    elif params is not None:
        assert abs(out["params"] / 1e6 - params) < 0.06        # SURVEY quotes these two in millions (285.7 M, 473 M)
    if cfg["wavelet_transform"]:
        assert out["sample_size"] == 128 and out["in_channels"] == 4 * cfg["image_channels"]


SCRIPT_STUBS = """
calls = {}
class _Loader:
    def __init__(self, **kw): calls["loader_kw"] = sorted(kw)
    def get_train_loader(self): return [0] * 7
    def get_val_loader(self): return [0] * 3
    def get_test_loader(self): return [0] * 3
def _mod(name, **attrs):
    m = types.ModuleType(name); m.__dict__.update(attrs); sys.modules[name] = m; return m
_mod("dataset")
_mod("dataset.chexpert", CheXpertDataLoader=_Loader)
_mod("dataset.cifar10", CIFAR10DataLoader=_Loader)
_mod("diffusers")
_mod("diffusers.optimization", get_cosine_schedule_with_warmup=lambda opt, num_warmup_steps, num_training_steps:
     calls.setdefault("sched", (num_warmup_steps, num_training_steps)))
_mod("accelerate", utils=types.SimpleNamespace(set_seed=lambda s: torch.manual_seed(s)))
from diffusion.diffusion_classifier import DiffusionClassifier
def _inference(self, **kw):          # no GPU in this container: record the call the script makes (GPU test runs it for real)
    calls["inference_kw"] = sorted(kw)
    calls["backbone"] = type(self.model).__module__
    calls["metrics"] = [type(m).__module__ + "." + type(m).__name__ for m in kw["metrics"]]
    calls["params"] = count(self.model)
    calls["classification"] = kw["classification"]
    return [{"accuracy": torch.tensor(1.0, device="cpu")}], None, None
DiffusionClassifier.inference = _inference
"""


@pytest.mark.parametrize("script,cfg,params", [
    ("experiments/chexpert-unet/inference.py",
     dict(image_size=256, image_channels=3, wavelet_transform=False, classes=2, batch_size=2, num_workers=0, seed=0,
          data_path="/nowhere", learning_rate=1e-4, lr_warmup_steps=10, num_epochs=2, classification=True,
          checkpoint_folder="checkpoints"), None),
    # (experiments/cifar10/inference.py cannot run against the reference either: its InferenceConfig has no __getattr__
    #  and it passes ``unet=`` to a constructor whose parameter is ``backbone`` -- a stale script, not a boundary)
    ("experiments/chexpert-dit/inference.py",
     dict(image_size=256, image_channels=3, wavelet_transform=False, classes=2, batch_size=2, num_workers=0, seed=0,
          data_path="/nowhere", learning_rate=1e-4, lr_warmup_steps=10, num_epochs=2, classification=True,
          checkpoint_folder="checkpoints", patch_size=4, encoder_type="DiT"), 147476784),
])
def test_reference_inference_script_runs_against_the_dropin(script, cfg, params, tmp_path):
    full = dict(project_root=str(tmp_path), experiment_dir="/exp", pred_param="eps", schedule="cosine", noise_d=64,
                cfg_w=0.0, ema_beta=0.999, ema_warmup=0, ema_update_freq=1, encoder_type="nn", n_stages=1,
                evaluation_per_stage=[2], n_keep_per_stage=[1], n_fast_classes=2, fast_classification=False,
                evaluation_batches=1)
    full.update(cfg)
    path = os.path.join(REF, script)
    body = SCRIPT_STUBS + f"""
import runpy
torch.set_default_device("meta")      # parameter containers only; nothing is computed here
g = runpy.run_path({path!r}, run_name="dropin_test")
g["main"]()
print(json.dumps(calls))
"""
    out = _run(body, env={"PROJECT_ROOT": REF, "TRAINING_CONFIG": json.dumps(full)})
    assert out["backbone"] == ("dcb200.dit" if "dit" in script else "dcb200.unet")
    assert all(m.startswith("dcb200.metrics.") for m in out["metrics"]) and len(out["metrics"]) >= 2
    assert {"val_dataloader", "metrics", "classification", "checkpoint_folder"} <= set(out["inference_kw"])
    assert out["classification"] is True
    if params:
        assert out["params"] == params
