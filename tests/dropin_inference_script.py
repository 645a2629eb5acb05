"""An experiment script written the way the reference's are (experiments/chexpert-unet/inference.py:1-22, 99-168): project
imports by the reference's module paths, a TRAINING_CONFIG JSON blob read through a __getattr__ config, the inline
(256, 512, 768) CheXpert U-Net, ``DiffusionClassifier(backbone=, config=).inference(...)`` with the four metrics.  Run by
tests/test_gpu_h_dropin.py with ``diffusion-classifier_b200/dropin`` first on PYTHONPATH: every import below resolves
to dcb200.  (The dataset module -- out of scope -- is replaced by a synthetic loader; accelerate / diffusers.optimization are
not needed because dcb200's ``inference`` loads the checkpoint itself.)"""
import json
import os
import sys

from nets.unet import UNetCondition2D
from diffusion.diffusion_classifier import DiffusionClassifier
from utils.metrics import Accuracy, F1, Precision, Recall
from utils.wavelet import wavelet_dec_2, wavelet_enc_2  # noqa: F401

import torch


class TrainingConfig:
    def __init__(self):
        self.config = json.loads(os.environ["TRAINING_CONFIG"])
        self.experiment_path = os.path.join(f"{self.config['project_root']}{self.config['experiment_dir']}")

    def __getattr__(self, name):
        return self.config.get(name)


def build(config):
    unet = UNetCondition2D(
        sample_size=config.image_size if not config.wavelet_transform else config.image_size // 2,
        in_channels=config.image_channels if not config.wavelet_transform else 4 * config.image_channels,
        out_channels=config.image_channels if not config.wavelet_transform else 4 * config.image_channels,
        layers_per_block=2,
        block_out_channels=(256, 512, 768),
        down_block_types=("DownBlock2D", "DownBlock2D", "CrossAttnDownBlock2D"),
        up_block_types=("CrossAttnUpBlock2D", "UpBlock2D", "UpBlock2D"),
        mid_block_type="UNetMidBlock2DCrossAttn",
        encoder_hid_dim=256,
        encoder_hid_dim_type='text_proj',
        cross_attention_dim=256,
    )
    return DiffusionClassifier(backbone=unet, config=config)


def loader(config, n):
    g = torch.Generator().manual_seed(5)
    S, C = config.image_size, config.image_channels
    return [{"images": torch.rand(config.batch_size, C, S, S, generator=g) * 2 - 1,
             "prompt": torch.randint(0, config.classes, (config.batch_size,), generator=g)} for _ in range(n)]


def main():
    config = TrainingConfig()
    torch.manual_seed(config.seed)
    if sys.argv[1] == "write":          # stands in for the training run that produced the checkpoint
        dc = build(config)
        with torch.no_grad():
            dc.encoder.weight.mul_(40.0)
            for p in dc.ema.ema_model.parameters():
                p.mul_(1.01)            # EMA weights differ from the online ones: inference must score with the EMA copy
        dc.save_checkpoint(os.path.join(config.experiment_path, config.checkpoint_folder), epoch=1)
        return
    torch.manual_seed(config.seed + 1)  # a different init: the weights must come from the checkpoint
    diffusion_classifier = build(config)
    metrics = [Accuracy("accuracy"), F1("f1"), Precision("precision"), Recall("recall")]
    torch.manual_seed(11)
    metric_output, samples, batches = diffusion_classifier.inference(
        train_dataloader=None, val_dataloader=loader(config, 3), optimizer=None, lr_scheduler=None, metrics=metrics,
        plot_function=None, classification=config.classification, checkpoint_folder=config.checkpoint_folder)
    out = [{k: round(float(v), 4) for k, v in d.items()} for d in metric_output]
    import dcb200
    print(json.dumps({"metrics": out, "labels": torch.cat(samples).tolist(),
                      "truth": torch.cat([b["prompt"] for b in batches]).tolist(),
                      "classes": [type(diffusion_classifier).__module__, type(diffusion_classifier.model).__module__,
                                  type(metrics[0]).__module__], "launches": dcb200.launch_count(),
                      "errors": diffusion_classifier.last_errors.mean(2).tolist()}))


if __name__ == "__main__":
    main()
