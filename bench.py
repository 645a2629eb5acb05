#!/usr/bin/env python
"""bench.py -- headline benchmark of the ELBO-classification hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload unet128|cifar|...]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...      (one rank per GPU, NCCL)

A "step" is one DiffusionClassifier.classify() pass over one batch of synthetic images of the workload's shape
(default: BASELINE configs[1], unet-128, 2 classes x 100 timesteps, bf16).  Prints ONE JSON line (rank 0).
  value     evals/s with the images already resident in HBM (CUDA events, max over ranks)
  e2e       same metric through the public API from pinned HOST images, H2D + label D2H inside the timed region
  roofline  tcgen05 implicit-GEMM kernel: algorithmic FLOPs of its launches / their summed CUDA-event durations
  cpu_baseline  the oracle port (oracle/: reference loop + restated diffusers U-Net, fp32) on the host cores
--impl reference times that CPU port alone (the reference's own implementation is pure Python over diffusers,
which is not installable offline).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "diffusion-classifier_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

WORKLOADS = {
    # name: (arch key in dcb200/configs.py, classes, timesteps, GFLOP/eval, images per GPU per step)
    "unet128": ("UNET128", 2, 100, 175.79, 4),
    "cifar": ("CIFAR_UNET", 10, 32, 10.454, 16),
    "dit": ("DIT_B4_256", 2, 250, 1315.0, 1),    # BASELINE configs[3]: CheXpert-256 DiT-B/4 (attention + adaLN kernels)
    "unet256": ("UNET256", 2, 250, 530.10, 1),   # BASELINE configs[2]: CheXpert 256x256 U-Net, shifted cosine schedule
    # BASELINE configs[4]: 10-channel 256x256 pixels -> GPU Haar DWT (/2) -> 40x128x128 wavelet-domain U-Net, full sweep
    "ipmsa": ("IPMSA5_DWT_UNET", 2, 1000, 189.96, 1),
}


def build_workload(name):
    from dcb200 import configs
    arch = getattr(configs, WORKLOADS[name][0])
    _, classes, T, gflop, ipg = WORKLOADS[name]
    S = arch["sample_size"]
    cfg = configs.classify_config(classes=classes, evaluation_per_stage=[T], n_stages=1, n_keep_per_stage=[1], noise_d=S,
                           image_size=S, schedule="cosine", pred_param="eps")
    if name == "dit":
        cfg.encoder_type = "DiT"
    if name == "unet256":
        cfg.schedule, cfg.noise_d = "shifted_cosine", 64
    if name == "ipmsa":
        cfg.wavelet_transform = True
    return arch, cfg, classes, T, gflop, ipg


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except (OSError, ValueError):
        return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0}, "fallback"


# ---------------------------------------------------------------------------------------------------------------
def cpu_port_run(arch, cfg, n_img, n_t, threads, seed=0):
    """the oracle port on the host: restated classify loop (oracle/loop.py) around the restated fp32 U-Net."""
    from oracle import diffusers_restated as dr
    from oracle import loop
    torch.set_num_threads(threads)
    torch.manual_seed(seed)
    if "patch_size" in arch:
        net, enc = dr.DiTTransformer2DModel(**arch).eval(), None
    else:
        net = dr.UNet2DConditionModel(**arch).eval()
        enc = torch.nn.Embedding(cfg.classes + 1, arch["encoder_hid_dim"])
    import copy
    c2 = copy.deepcopy(cfg)
    c2.evaluation_per_stage = [n_t]
    S, C = arch["sample_size"], arch["in_channels"]
    x = torch.rand(n_img, C, S, S) * 2 - 1

    class Den(torch.nn.Module):
        def forward(self, x, noise_labels, encoder_hidden_states):
            return net(x, noise_labels, encoder_hidden_states)[0]

    den = Den()
    t0 = time.perf_counter()
    loop.classify_oracle(den, enc, c2, x)
    dt = time.perf_counter() - t0
    return n_img * n_t * cfg.classes / dt, dt


def gpu_eager_port_run(arch, cfg, n_img, n_t, dev, autocast, seed=0):
    """the same oracle port (reference loop around the restated diffusers denoiser) in plain torch eager ON THE GPU: what the
    reference's own code path costs on this B200 (cuDNN / cuBLAS kernels, one forward of batch BS per class and timestep).
    A second reference line next to the CPU one; test infrastructure, never part of the product path."""
    from oracle import diffusers_restated as dr
    from oracle import loop
    import copy
    torch.manual_seed(seed)
    if "patch_size" in arch:
        net, enc = dr.DiTTransformer2DModel(**arch).eval().to(dev), None
    else:
        net = dr.UNet2DConditionModel(**arch).eval().to(dev)
        enc = torch.nn.Embedding(cfg.classes + 1, arch["encoder_hid_dim"]).to(dev)
    c2 = copy.deepcopy(cfg)
    c2.evaluation_per_stage = [n_t]
    S, C = arch["sample_size"], arch["in_channels"]
    x = (torch.rand(n_img, C, S, S) * 2 - 1).to(dev)

    class Den(torch.nn.Module):
        def forward(self, x, noise_labels, encoder_hidden_states):
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                return net(x, noise_labels, encoder_hidden_states)[0].float()

    den = Den()
    with torch.no_grad():
        loop.classify_oracle(den, enc, c2, x[:1])          # warm-up (cuDNN autotune, allocator)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        loop.classify_oracle(den, enc, c2, x)
        e1.record()
        torch.cuda.synchronize()
    dt = e0.elapsed_time(e1) / 1e3
    del net
    torch.cuda.empty_cache()
    return n_img * n_t * cfg.classes / dt, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    arch, cfg, classes, T, gflop, ipg = build_workload(args.workload)
    threads = len(os.sched_getaffinity(0))
    n_t = {"unet128": 16, "cifar": 8, "ipmsa": 8}.get(args.workload, 2)
    for _ in range(args.warmup):
        cpu_port_run(arch, cfg, 1, 1, threads)
    vals, times = [], []
    for _ in range(args.steps):
        v, dt = cpu_port_run(arch, cfg, 1, n_t, threads)
        vals.append(v)
        times.append(dt)
    val = sum(vals) / len(vals)
    sample = f"1 image x {n_t} timesteps x {classes} classes ({n_t * classes} denoiser evals) per step, fp32, torch CPU"
    line = {
        "impl": "reference", "metric": "denoiser evals/sec", "value": val, "unit": "evals/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "images_per_sec": val / (classes * T),
        "config": {"workload": workload_name(args.workload, classes, T), "sample": sample},
        "cpu_baseline": {"value": val, "unit": "evals/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_name(w, classes, T):
    return {"unet128": f"unet-128 class-conditional U-Net (models/unet-128.py) ELBO scoring, {classes} classes x {T} "
                       f"timesteps, 3x128x128",
            "cifar": f"CIFAR-10 32x32 class-conditional U-Net (experiments/cifar10) ELBO classification, {classes} "
                     f"classes x {T} timesteps",
            "dit": f"CheXpert 256x256 DiT-B/4 (models/chexpert-256-dit-b4) ELBO classification, {classes} classes x {T} "
                   f"timesteps",
            "unet256": f"CheXpert 256x256 U-Net (models/unet-256.py) healthy/sick ELBO classification, {classes} classes x "
                       f"{T} timesteps, 3x256x256, shifted cosine schedule",
            "ipmsa": f"IPMSA 5-channel-pair DWT U-Net (models/ipmsa-5-dwt-unet.py): 10x256x256 pixels -> Haar DWT on the "
                     f"GPU -> 40x128x128 wavelet-domain ELBO classification, {classes} classes x {T} timesteps"}[w]


# ---------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    import dcb200
    from dcb200 import engine as E
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # keep stdout to the ONE JSON line: NCCL prints its version banner on stdout when the communicator is created
        # (NCCL_DEBUG=VERSION on the boxes), so file descriptor 1 points at stderr while that happens
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    arch, cfg, classes, T, gflop, ipg = build_workload(args.workload)
    if args.images:
        ipg = args.images
    # weak scaling (default): every rank owns 1/world of the (image x timestep) units of a world-x larger batch;
    # strong scaling (--scaling strong): ONE batch of `ipg` images, its units split over the ranks (the latency case)
    BS = ipg * (world if args.scaling == "weak" else 1)
    cfg.dcb_shard = "timestep"
    if args.max_batch:
        cfg.dcb_max_batch = args.max_batch
    torch.manual_seed(0)
    net = (dcb200.DiT if args.workload == "dit" else dcb200.UNetCondition2D)(**arch)
    dc = dcb200.DiffusionClassifier(net, cfg).to(dev).eval()
    S, C = arch["sample_size"], arch["in_channels"]
    g = torch.Generator().manual_seed(0)
    if args.workload == "ipmsa":   # pixel-space input; the wavelet transform (utils/wavelet.py, /2 as in
        x_host = (torch.rand(BS, C // 4, 2 * S, 2 * S, generator=g) * 2 - 1).pin_memory()   # experiments/ipmsa/inference.py:153-155)
        pre = lambda v: dcb200.wavelet_dec_2(v, 0.5)                                        # runs on the GPU inside every step
    else:
        x_host = (torch.rand(BS, C, S, S, generator=g) * 2 - 1).pin_memory()
        pre = lambda v: v
    x_dev = x_host.to(dev)
    evals_per_step = BS * classes * T

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.path != "classify":
        return run_next_row(args, dc, cfg, x_host, x_dev, pre, classes, T, world, rank, dev, barrier)

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def step_resident():
        torch.manual_seed(1234)  # same t stream on every rank (CPU generator, as diffusion_classifier.py:688)
        return dc.classify(pre(x_dev))

    def step_e2e():
        torch.manual_seed(1234)
        labels = dc.classify(pre(x_host.to(dev, non_blocking=True)))
        return labels.cpu()

    for _ in range(max(args.warmup, 3)):
        step_resident()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n0 = dcb200.launch_count()
    ms = timed(step_resident, args.steps)
    launches = dcb200.launch_count() - n0
    clocks = sampler.stop() if rank == 0 else None
    value = evals_per_step * args.steps / (ms / 1e3)

    step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    e2e_value = evals_per_step * args.steps / (ms_e2e / 1e3)

    # roofline of the dominant kernel: CUDA-event pairs around every tcgen05 GEMM launch of one instrumented pass
    cfg.dcb_cuda_graph = False  # event pairs need eager launches
    E.PROFILE = prof = E.GemmProfile()
    barrier()
    for _ in range(min(2, args.steps)):
        step_resident()
    barrier()
    E.PROFILE = None
    gemm_ms, gemm_flops, n_gemm = prof.totals()
    # FLOPs of the same launches on the REFERENCE's graph (SURVEY 8d "algorithmic"): one untimed pass with the exact
    # work-saving rewrites switched off (per-class prefix recomputation, unfolded Upsample2D); only its FLOP count is used
    E.PROFILE = prof_ref = E.GemmProfile()
    fold0, E.FOLD_UPSAMPLE, cfg.dcb_share_prefix = E.FOLD_UPSAMPLE, False, False
    step_resident()
    barrier()
    E.PROFILE, E.FOLD_UPSAMPLE, cfg.dcb_share_prefix = None, fold0, None
    _, ref_flops, _ = prof_ref.totals()
    cfg.dcb_cuda_graph = None
    peaks, peak_src = measured_peaks()
    peak = peaks["bf16_tflops_sustained"]
    passes = min(2, args.steps)
    # the dominant kernel on the FLOPs its launches EXECUTE (2 M N K of every tcgen05 GEMM launch of the instrumented
    # passes) over their summed CUDA-event durations: this is the roofline fraction
    achieved = gemm_flops / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
    # the same launch time against the REFERENCE graph's conv / linear FLOPs (per-class prefix recomputed,
    # Upsample2D unfolded): an effective rate that credits the exact work-saving rewrites -- never a fraction of peak
    algorithmic = ref_flops * passes / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
    gn_ms, gn_bytes, n_gn = prof.gn_totals()
    hbm_peak = peaks["hbm_gbs"]
    gn_gbs = gn_bytes / (gn_ms / 1e3) / 1e9 if gn_ms > 0 else 0.0
    roofline = {
        "bound": "tensor", "kernel": "tcgen05 GEMM kernels (gemm_tc, gemm_tc2, gemm_tc2x, gemm_tc3: implicit-GEMM convs and "
                                     "linears)", "achieved": achieved,
        "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": None,
        "peak_source": f"{peak_src} bf16_tflops_sustained",
        "achieved_basis": "executed FLOPs (2 M N K of every tcgen05 GEMM launch) / summed CUDA-event durations of those "
                          "launches, eager launches on the bench workload; since round 2 the full-resolution conv launches also "
                          "apply the GroupNorm + SiLU of their operand (gemm_tc2x), so their time contains what used to be the "
                          "separate gn_apply pass",
        "effective_tflops_on_reference_graph": algorithmic,
        "effective_speedup_vs_reference_flops": (ref_flops * passes) / gemm_flops if gemm_flops else None,
        "launches_per_step": n_gemm // passes, "avg_launch_ms": gemm_ms / max(n_gemm, 1),
        "kernel_share_of_step": (gemm_ms / passes) / (ms / args.steps),
        "executed_gflop_per_eval": gemm_flops / passes / (evals_per_step / world) / 1e9,
        "reference_graph_gemm_gflop_per_eval": ref_flops / (evals_per_step / world) / 1e9,
        "reference_gflop_per_eval": gflop,
        "whole_step_frac_of_peak": (gemm_flops / passes) / (ms / args.steps / 1e3) / 1e12 / peak,
        "whole_step_frac_of_peak_on_reference_flops":
            gflop * 1e9 * (evals_per_step / world) / (ms / args.steps / 1e3) / 1e12 / peak,
        # second kernel by share of the step: the streaming GroupNorm(+SiLU) apply pass, HBM bound
        "secondary": None if n_gn == 0 else {
            "kernel": "gn_apply_kernel (GroupNorm + SiLU apply, streaming)", "bound": "hbm", "achieved": gn_gbs,
            "peak": hbm_peak, "unit": "GB/s", "frac": gn_gbs / hbm_peak, "peak_source": f"{peak_src} hbm_gbs",
            "achieved_basis": "one read + one write of every normalised tensor / summed CUDA-event durations",
            "launches_per_step": n_gn // passes, "kernel_share_of_step": (gn_ms / passes) / (ms / args.steps)},
    }
    at_ms, at_flops, n_at = prof.attn_totals()
    if at_ms > gn_ms and n_at:   # DiT: attention, not GroupNorm, is the second kernel of the step
        at_tf = at_flops / (at_ms / 1e3) / 1e12
        roofline["secondary"] = {
            "kernel": "flash_attn_tc_fast_kernel / flash_attn_tc_kernel (tcgen05 attention, incl. the norm pre-pass)",
            "bound": "tensor", "achieved": at_tf, "peak": peak, "unit": "TFLOP/s", "frac": at_tf / peak,
            "peak_source": f"{peak_src} bf16_tflops_sustained",
            "achieved_basis": "4 B h N^2 d / summed CUDA-event durations of the attention calls; the softmax (one exponential "
                              "per score on the XU / FMA pipes) bounds this kernel below the tensor peak at head dim 64",
            "launches_per_step": n_at // passes, "kernel_share_of_step": (at_ms / passes) / (ms / args.steps)}
    tr = os.path.join(ROOT, "profiles", "r02_traffic.json")   # per-launch DRAM bytes of the dominant kernel from the
    if os.path.exists(tr):                                     # committed `ncu --set full` capture (same workload)
        try:
            tj = json.load(open(tr))
            per_eval = tj.get(args.workload + "_dram_bytes_per_eval")
            if per_eval:   # ncu DRAM bytes of all GEMM launches per eval -> per launch at this run's launch size
                roofline["traffic"] = per_eval * (evals_per_step / world) / max(n_gemm // passes, 1)
                roofline["traffic_source"] = "ncu dram__bytes_read+write of every tcgen05 GEMM launch of one pass (profiles/r02_traffic.json), scaled to this launch size"
        except (OSError, ValueError):
            pass

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = len(os.sched_getaffinity(0))
        n_t = {"unet128": 32, "cifar": 8, "ipmsa": 16}.get(args.workload, 4)
        cpu_port_run(arch, cfg, 1, 1, threads)
        v, dt = cpu_port_run(arch, cfg, 1, n_t, threads)
        cpu = {"value": v, "unit": "evals/s", "cores": threads, "kind": "port", "seconds": dt,
               "sample": f"1 image x {n_t} timesteps x {classes} classes ({n_t * classes} evals) of the same workload, "
                         f"fp32 oracle port (reference loop + restated diffusers denoiser) on torch CPU"}

    eager = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            # an eager-favourable batch: the reference issues one forward of batch BS per (class, timestep); at BS = 4 torch
            # eager is launch-bound, so it is timed at a batch that fills the GPU
            n_img = {"unet128": 32, "cifar": 256, "ipmsa": 16, "unet256": 8, "dit": 8}.get(args.workload, 8)
            n_t = 2
            rows = {}
            for tag, ac in (("fp32 (torch default: TF32 cuDNN convs, fp32 matmuls)", False), ("bf16 autocast", True)):
                v, dt = gpu_eager_port_run(arch, cfg, n_img, n_t, dev, ac)
                rows[tag] = {"value": v, "seconds": dt}
            eager = {"unit": "evals/s", "kind": "port", "by_dtype": rows,
                     "sample": f"{n_img} images x {n_t} timesteps x {classes} classes, one forward of batch {n_img} per (class, "
                               f"timestep) as diffusion_classifier.py:686-704 issues them; oracle port in torch eager on cuda:0"}
        except Exception as ex:  # the checker must never take the bench line down
            eager = {"error": repr(ex)[:200]}

    if rank == 0:
        line = {
            "metric": "denoiser evals/sec", "value": value, "unit": "evals/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "images_per_sec": value / (classes * T),
            "config": {"workload": workload_name(args.workload, classes, T), "images_per_step": BS,
                       "evals_per_step": evals_per_step, "shard": "(image x timestep) units over ranks + 1 all-reduce",
                       "eps": "in-kernel Philox", "cuda_graph": True, "weights": "random init (torch default, seed 0)",
                       "shared_prefix": "layers ahead of the first cross-attention run once per (image, timestep)",
                       "l2": "activation working set per launch sequence is GBs (>> 126 MB L2); no flush needed"},
            "e2e": {"value": e2e_value, "unit": "evals/s", "h2d_bytes_per_step": x_host.numel() * 4,
                    "d2h_bytes_per_step": BS * 8, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "torch_eager_gpu_baseline": eager,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_next_row(args, dc, cfg, x_host, x_dev, pre, classes, T, world, rank, dev, barrier):
    """SURVEY 8(f) rows measured like the hot path: CUDA events around K steps, inputs resident (value) and from pinned
    host memory with the result read back (e2e).  sample: DDPM + CFG, 2 evaluations per step and image (+1 final step);
    loss: one forward per image; evaluate: classify through DiffusionClassifier.evaluate with a host-side loader
    (prefetched H2D) and GPU-side metrics."""
    import dcb200
    BS = x_dev.shape[0]
    g = torch.Generator().manual_seed(1)
    text_host = torch.randint(0, classes, (BS,), generator=g).pin_memory()
    text_dev = text_host.to(dev)
    if args.path == "sample":
        cfg.sampling_steps, cfg.cfg_w = 20, 1.0
        dc.cfg_w = 1.0
        evals = BS * 2 * (cfg.sampling_steps + 1)
        res = lambda: dc.sample(pre(x_dev), text_dev)
        e2e = lambda: dc.sample(pre(x_host.to(dev, non_blocking=True)), text_host.to(dev, non_blocking=True)).cpu()
        d2h = x_host.numel() * 4
        what = f"DiffusionClassifier.sample: DDPM, {cfg.sampling_steps} steps + final, cfg_w=1 (cond + uncond folded into the batch)"
    elif args.path == "loss":
        evals = BS
        res = lambda: dc.loss(pre(x_dev), text_dev)
        e2e = lambda: dc.loss(pre(x_host.to(dev, non_blocking=True)), text_host.to(dev, non_blocking=True)).item()
        d2h = 4
        what = "DiffusionClassifier.loss forward (min-SNR weighted eps-MSE of the online model)"
    else:
        nb = 3
        loader = [{"images": x_host.clone(), "prompt": text_host.clone()} for _ in range(nb)]
        loader_dev = [{"images": pre(x_dev), "prompt": text_dev} for _ in range(nb)]
        evals = nb * BS * classes * T
        mets = [dcb200.metrics.Accuracy("accuracy"), dcb200.metrics.F1("f1")]
        for m in mets:
            m.set_device(dev)

        def run(ld):
            torch.manual_seed(1234)
            _, _, ms_ = dc.evaluate(ld, metrics=mets, classification=True)
            return ms_

        res = lambda: run(loader_dev)
        e2e = lambda: float(run(loader)[0].get_output()["accuracy"])
        d2h = 16
        what = f"DiffusionClassifier.evaluate: {nb} host batches of {BS} images, next batch's H2D overlapped, GPU-side Accuracy/F1"

    def timed(fn):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            import torch.distributed as dist
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(max(args.warmup, 3)):
        res()
    sampler = ClockSampler(dev.index or 0)
    if rank == 0:
        sampler.start()
    n0 = dcb200.launch_count()
    ms = timed(res)
    launches = dcb200.launch_count() - n0
    clocks = sampler.stop() if rank == 0 else None
    e2e()
    ms2 = timed(e2e)
    if rank == 0:
        wl = workload_name(args.workload, classes, T)
        # replicas: every rank runs the same step on its own images (no collective on these paths)
        print(json.dumps({
            "metric": "denoiser evals/sec", "value": world * evals * args.steps / (ms / 1e3), "unit": "evals/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{what}; network and image shape of: {wl}", "path": args.path,
                       "images_per_step": BS * world, "evals_per_step": evals * world, "replicas": world},
            "e2e": {"value": world * evals * args.steps / (ms2 / 1e3), "unit": "evals/s",
                    "h2d_bytes_per_step": x_host.numel() * 4 + BS * 8, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms2 / args.steps},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": None, "cpu_baseline": None}))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="unet128", choices=list(WORKLOADS))
    ap.add_argument("--images", type=int, default=0, help="images per GPU per step")
    ap.add_argument("--max-batch", type=int, default=0, help="denoiser samples per launch sequence")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: images per step = --images x N (default); strong: one fixed batch split over the N ranks")
    ap.add_argument("--path", default="classify", choices=["classify", "sample", "loss", "evaluate"],
                    help="which caller of the denoiser kernels a step is (SURVEY 8f rows; default = the hot path)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
