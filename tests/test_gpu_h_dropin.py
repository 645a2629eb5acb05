"""Drop-in path on the GPU: an experiment script in the reference's style (tests/dropin_inference_script.py, modelled on
experiments/chexpert-unet/inference.py:99-168 -- the reference tree itself is not on the GPU box) runs in a fresh interpreter
with ``dropin/`` first on PYTHONPATH: checkpoint written, loaded by ``inference()``, two batches classified with the
inline (256, 512, 768) CheXpert U-Net (attention at 1 024 tokens, head dim 96), metrics synced and printed.  The labels
must equal a direct dcb200 classify of the same batches with the checkpoint's EMA weights."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "diffusion-classifier_b200")


def test_reference_style_inference_script_through_dropin(dev, tmp_path):
    cfg = dict(project_root=str(tmp_path), experiment_dir="/exp", pred_param="eps", schedule="shifted_cosine", noise_d=64,
               image_size=128, image_channels=3, wavelet_transform=False, cfg_w=0.0, ema_beta=0.999, ema_warmup=0,
               ema_update_freq=1, encoder_type="nn", classes=2, n_stages=1, evaluation_per_stage=[3],
               n_keep_per_stage=[1], n_fast_classes=2, fast_classification=False, evaluation_batches=1, batch_size=2,
               seed=0, classification=True, checkpoint_folder="checkpoints")
    env = dict(os.environ, TRAINING_CONFIG=json.dumps(cfg),
               PYTHONPATH=os.pathsep.join([os.path.join(PKG, "dropin"), PKG]))
    script = os.path.join(ROOT, "tests", "dropin_inference_script.py")
    for mode in ("write", "infer"):
        r = subprocess.run([sys.executable, script, mode], capture_output=True, text=True, timeout=600, env=env)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    out = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert out["classes"] == ["dcb200.classifier", "dcb200.unet", "dcb200.metrics"] and out["launches"] > 0
    assert len(out["labels"]) == 4                         # evaluation_batches = 1 is an inclusive stop index (:574)
    acc = sum(int(a == b) for a, b in zip(out["labels"], out["truth"])) / 4
    assert abs(out["metrics"][0]["accuracy"] - acc) < 1e-4 and {"f1", "precision", "recall"} <= {
        k for d in out["metrics"] for k in d}
    # the same batches, scored directly with the checkpoint's EMA weights
    sys.path[:0] = [os.path.join(ROOT, "tests")]
    import dcb200
    from dcb200 import configs
    # (build the network here without the dropin path: same class, imported as dcb200.*)
    c = configs.Config(**cfg, experiment_path=str(tmp_path) + "/exp")
    net = dcb200.UNetCondition2D(
        sample_size=128, in_channels=3, out_channels=3, layers_per_block=2, block_out_channels=(256, 512, 768),
        down_block_types=("DownBlock2D", "DownBlock2D", "CrossAttnDownBlock2D"),
        up_block_types=("CrossAttnUpBlock2D", "UpBlock2D", "UpBlock2D"), mid_block_type="UNetMidBlock2DCrossAttn",
        encoder_hid_dim=256, encoder_hid_dim_type="text_proj", cross_attention_dim=256)
    dc = dcb200.DiffusionClassifier(net, c)
    dc.load_checkpoint(os.path.join(str(tmp_path) + "/exp", "checkpoints"))
    assert not torch.equal(dc.ema.ema_model.conv_in.weight, dc.model.conv_in.weight)
    dc = dc.to(dev).eval()
    g = torch.Generator().manual_seed(5)
    batches = [{"images": torch.rand(2, 3, 128, 128, generator=g) * 2 - 1, "prompt": torch.randint(0, 2, (2,), generator=g)}
               for _ in range(3)]
    torch.manual_seed(11)
    labels = torch.cat([dc.classify(b["images"].to(dev)) for b in batches[:2]]).tolist()
    assert labels == out["labels"]
    assert out["truth"] == torch.cat([b["prompt"] for b in batches[:2]]).tolist()
