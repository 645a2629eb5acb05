"""ONE eager classify pass of a bench workload between cudaProfilerStart / Stop (3 warm-up passes before it), for
`ncu --profile-from-start off`:   python tools/one_pass.py [workload] [images] [max_batch]
Prints the pass's CUDA-event time, the number of libdcb200 launches in it and the evals it scored."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "diffusion-classifier_b200")):
    sys.path.insert(0, p)
import torch
import bench, dcb200

wl = sys.argv[1] if len(sys.argv) > 1 else "unet128"
images = int(sys.argv[2]) if len(sys.argv) > 2 else 1
arch, cfg, classes, T, gflop, ipg = bench.build_workload(wl)
if len(sys.argv) > 3 and int(sys.argv[3]):
    cfg.dcb_max_batch = int(sys.argv[3])
cfg.dcb_cuda_graph = False
dev = torch.device("cuda:0")
torch.manual_seed(0)
dc = dcb200.DiffusionClassifier((dcb200.DiT if wl == "dit" else dcb200.UNetCondition2D)(**arch), cfg).to(dev).eval()
S, C = arch["sample_size"], arch["in_channels"]
if wl == "ipmsa":
    x = dcb200.wavelet_dec_2((torch.rand(images, C // 4, 2 * S, 2 * S) * 2 - 1).to(dev), 0.5)
else:
    x = (torch.rand(images, C, S, S) * 2 - 1).to(dev)
for _ in range(3):
    torch.manual_seed(1234)
    dc.classify(x)
torch.cuda.synchronize()
n0 = dcb200.launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.profiler.start()
e0.record()
torch.manual_seed(1234)
dc.classify(x)
e1.record()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(json.dumps({"workload": wl, "images": images, "evals": images * classes * T, "pass_ms": e0.elapsed_time(e1),
                  "dcb_launches": dcb200.launch_count() - n0, "max_batch": cfg.dcb_max_batch}))
