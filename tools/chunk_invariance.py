"""which knob breaks bit-identity of the unet-128 error table across launch-sequence sizes? (debug tool)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "diffusion-classifier_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import dcb200
from dcb200 import engine as E
from dcb200 import _lib as L
from helpers import UNET128, base_cfg
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = dcb200.UNetCondition2D(**UNET128)
cfg = base_cfg(classes=2, evaluation_per_stage=[6], noise_d=128, image_size=128)
dc = dcb200.DiffusionClassifier(net, cfg).to(dev).eval()
x = torch.rand(3, 3, 128, 128, device=dev, generator=torch.Generator(device=dev).manual_seed(1)) * 2 - 1


def run(mb, graph=False, share=None):
    cfg.dcb_max_batch, cfg.dcb_cuda_graph, cfg.dcb_share_prefix = mb, graph, share
    dc._eps_calls = 0
    torch.manual_seed(5)
    dc.classify(x)
    return dc.last_errors.clone()


base = run(0)
for name, env in (("default", {}), ("tile stats off", {"ts": False}), ("fold off", {"fold": False}),
                  ("fused small gn off", {"fsg": False})):
    E.USE_TILE_STATS = env.get("ts", True)
    E.FOLD_UPSAMPLE = env.get("fold", True)
    E.USE_FUSED_SMALL_GN = env.get("fsg", True)
    b0 = run(0)
    row = []
    for mb in (36, 18, 8, 2):
        t = run(mb)
        row.append(f"mb={mb}: {'==' if torch.equal(t, b0) else 'DIFF %.1e' % float(((t - b0).abs() / b0).max())}")
    t = run(0, share=False)
    row.append(f"share off: {'==' if torch.equal(t, b0) else 'DIFF %.1e' % float(((t - b0).abs() / b0).max())}")
    print(f"{name:22s}", "  ".join(row), flush=True)
for knob in ("NO_TC2", "TC2_NO_HALO", "TC2_NO_YHALO", "NO_TC2_MSE"):
    E.USE_TILE_STATS = E.FOLD_UPSAMPLE = E.USE_FUSED_SMALL_GN = True
    kn = L.knob(knob); kn.__enter__()
    b0 = run(0)
    row = [f"vs default: {'==' if torch.equal(b0, base) else 'DIFF %.1e' % float(((b0 - base).abs() / base).max())}"]
    for mb in (8, 2):
        t = run(mb)
        row.append(f"mb={mb}: {'==' if torch.equal(t, b0) else 'DIFF %.1e' % float(((t - b0).abs() / b0).max())}")
    kn.__exit__()
    print(f"{knob:22s}", "  ".join(row), flush=True)
