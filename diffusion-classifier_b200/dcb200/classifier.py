"""DiffusionClassifier -- drop-in for the reference's hot path, diffusion/diffusion_classifier.py:17-161,657-725.

Same constructor (``DiffusionClassifier(backbone, config)``), same attributes (``model, ema, encoder, null_token,
schedule, pred_param, config``) and the same ``classify(x, text=None, fast=False) -> LongTensor[BS]`` contract,
including its assertions, its CPU-generator draw of t (:688), its ``errors`` table initialised to +inf (:669)
and its per-stage ``mean -> topk(largest=False)`` pruning (:718-721).

What changes is how the work is issued (B200-first):
  * the (image b, timestep j, alive class c) triple loop is folded into the denoiser's batch axis: one launch
    sequence scores ``U`` (b, j) units x ``n_alive`` classes, sharing z_t / eps across the classes of a unit;
  * q_sample is a fused prologue kernel that also stages the first layer's operand (dcb_prologue);
  * the eps-MSE is the epilogue of the last GEMM (predicted noise never reaches HBM);
  * the class token's cross-attention collapses to a per-class bias table computed once per call;
  * with torch.distributed initialised and ``config.dcb_shard == 'timestep'`` the (b, j) units are sharded over
    ranks and the stage's error slab is combined by ONE all-reduce (NCCL over NVLink) before the top-k.
"""
from __future__ import annotations

import math
import os
import time

import torch
import torch.nn as nn

from . import _lib as L
from . import engine as E
from .dit import DiT
from .ema import EMA
from .unet import UNetCondition2D


def log(t, eps=1e-20):
    return torch.log(t.clamp(min=eps))


def shard_range(n_units, rank, world):
    """Contiguous slice of the stage's flattened (timestep j, image b) units owned by ``rank``."""
    return (n_units * rank) // world, (n_units * (rank + 1)) // world


def combine_stage_errors(errors, slab, classes, start, end, dist=None):
    """Merge one stage's per-rank error slab [BS, classes, nj] (zeros where a rank owns nothing) into the
    reference-shaped ``errors`` table: ONE sum all-reduce, then only the (image, alive class) entries are written so
    pruned classes keep +inf exactly as in diffusion_classifier.py:669,713-721.  Adding zeros is exact, hence labels
    and sums are bit-identical for every world size."""
    if dist is not None:
        dist.all_reduce(slab)
    alive = torch.zeros(errors.shape[0], errors.shape[1], dtype=torch.bool, device=errors.device)
    alive.scatter_(1, classes, True)
    errors[:, :, start:end] = torch.where(alive.unsqueeze(-1), slab, errors[:, :, start:end])
    return errors


def sync_from_rank0(dist, t, dev):
    """``t`` as rank 0 holds it, on ``dev``.  With ``dcb_shard == 'timestep'`` every rank scores a slice of the SAME
    (image, timestep, class) table, so everything that defines the table -- the timesteps drawn from the CPU generator
    (:688), the fast-mode candidate classes (:671-677) and the Philox seed -- must be rank 0's: ranks seeded differently
    (the usual seed + rank set-up) would otherwise fill different class columns and the all-reduced table would hold
    zeros where the reference has errors.  (The reference never needed this: accelerate shards images, :613-617.)"""
    tt = t.to(dev).contiguous()
    if dist is not None:
        dist.broadcast(tt, 0)
    return tt


class _GraphCache:
    """LRU of captured launch sequences keyed by (kind, pack generation, shapes).  Entries of a superseded pack generation of
    the same network are dropped first (their weights are gone); capacity grows with the number of pruning stages (each
    stage contributes a full-chunk and a tail-chunk shape) so one classify() call never evicts its own graphs."""

    def __init__(self):
        import collections
        self.d = collections.OrderedDict()
        self.cap = 8

    def get(self, key):
        v = self.d.get(key)
        if v is not None:
            self.d.move_to_end(key)
        return v

    def put(self, key, value, net_id, gen):
        for k in [k for k, v in self.d.items() if v["net_id"] == net_id and v["gen"] != gen]:
            del self.d[k]
        while len(self.d) >= self.cap:
            self.d.popitem(last=False)
        self.d[key] = dict(obj=value, net_id=net_id, gen=gen)
        return value

    def __len__(self):
        return len(self.d)


class _GraphedDenoiser:
    """One CUDA graph per (network, chunk shape): the ~500 kernel launches of a denoiser pass + fused eps-MSE are
    captured once and replayed; the prologue writes straight into the graph's fixed input buffers.  (CUDA streams and
    graphs instead of a tracing compiler -- the captured launches are exactly the eager ones.)"""
    _pool = None

    def __init__(self, net, ctx, pk, is_dit, U, nk, Cimg, H, W, patch, v_param, fused, share):
        dev = ctx.device
        rows = (H // patch) * (W // patch)
        S = U * nk
        self.share = share
        self.a_in = ctx.empty((U if share else S) * rows, pk.kpad_in)
        self.target = torch.empty(U * H * W * Cimg, device=dev, dtype=torch.float32)
        self.logsnr = torch.empty(U, device=dev, dtype=torch.float32)
        self.cls = torch.empty(S, device=dev, dtype=torch.int32)
        self.scale = torch.empty(S, device=dev, dtype=torch.float32) if v_param else None
        self.err = torch.empty(S, device=dev, dtype=torch.float32)
        self.table = None          # the graph's OWN copy of the cross-attention table (fixed address for its lifetime)
        self._table_call = None
        self.graph = None
        self.args = (net, ctx, pk, is_dit, U, nk, H, W, (patch * patch * Cimg) if is_dit else Cimg, fused)
        self.warm = 0

    def set_table(self, tb, call_id):
        """copy this call's class table into the graph's persistent buffer (once per call)"""
        if tb is None:
            return
        if self.table is None:
            self.table = torch.empty_like(tb)
        if self._table_call != call_id:
            self.table.copy_(tb)
            self._table_call = call_id

    def _run(self):
        net, ctx, pk, is_dit, U, nk, H, W, No, fused = self.args
        mse = dict(target=self.target, div=nk, ld=No, err=self.err, fused=fused, scale=self.scale)
        if is_dit:
            net.run(ctx, pk, self.a_in, self.logsnr, U, nk, self.cls, mse=mse)
        else:
            net.run(ctx, pk, self.a_in, self.logsnr, U, nk, H, W, self.table, xattn_idx=self.cls, mse=mse,
                    share_prefix=self.share)

    def launch(self):
        if self.graph is not None:
            self.graph.replay()
            L.lib().dcb_note_graph_replay(self.n_kernels)
            return
        if self.warm < 2:                # eager first (sets kernel attributes, fills the allocator / tensor-map caches)
            self._run()
            self.warm += 1
            return
        torch.cuda.synchronize()         # third use of this chunk shape: capture, then replay
        g = torch.cuda.CUDAGraph()
        if _GraphedDenoiser._pool is None:
            _GraphedDenoiser._pool = torch.cuda.graph_pool_handle()
        n0 = L.launch_count()
        with torch.cuda.graph(g, pool=_GraphedDenoiser._pool):
            self._run()
        self.n_kernels = L.launch_count() - n0
        L.lib().dcb_note_graph_replay(-self.n_kernels)   # captured launches did not execute
        self.graph = g
        self.graph.replay()
        L.lib().dcb_note_graph_replay(self.n_kernels)


class DiffusionClassifier(nn.Module):
    def __init__(self, backbone: nn.Module, config):
        super().__init__()
        self.config = config
        pred_param = self.config.pred_param
        assert pred_param in ['v', 'eps'], "Invalid prediction parameterization. Must be 'v' or 'eps'"
        self.pred_param = pred_param
        schedule = self.config.schedule
        assert schedule in ['cosine', 'shifted_cosine'], "Invalid schedule. Must be 'cosine' or 'shifted_cosine'"
        self.schedule = self.logsnr_schedule_cosine if schedule == 'cosine' else self.logsnr_schedule_cosine_shifted
        self.noise_d = self.config.noise_d
        self.image_d = self.config.image_size
        self.cfg_w = self.config.cfg_w
        assert isinstance(backbone, nn.Module), "Model must be an instance of torch.nn.Module."
        self.model = backbone
        self.ema = EMA(self.model, beta=config.ema_beta, update_after_step=config.ema_warmup,
                       update_every=config.ema_update_freq)
        self.encoder_type = self.config.encoder_type
        if self.encoder_type == 't5':
            raise NotImplementedError("the t5 text encoder is not on the classification hot path")
        elif self.encoder_type == 'nn':
            self.encoder = nn.Embedding(self.config.classes + 1, backbone.config.encoder_hid_dim)
            self.tokenizer = None
            self.null_token = self.config.classes
        elif self.encoder_type == 'DiT':
            self.tokenizer = None
            self.encoder = None
            self.null_token = self.config.classes
        self.last_errors = None  # [BS, classes, T] fp32 table of the most recent classify() call
        self._eps_calls = 0
        self._graphs = _GraphCache()
        self._mem_probe = None     # (time, bytes the device could still give us) of the last cudaMemGetInfo

    # ---- schedule (diffusion_classifier.py:119-161), evaluated exactly as the reference does ------------------
    def logsnr_schedule_cosine(self, t, logsnr_min=-15, logsnr_max=15):
        logsnr_max = logsnr_max + math.log(self.noise_d / self.image_d)
        logsnr_min = logsnr_min + math.log(self.noise_d / self.image_d)
        t_min = math.atan(math.exp(-0.5 * logsnr_max))
        t_max = math.atan(math.exp(-0.5 * logsnr_min))
        return -2 * log(torch.tan(t_min + t * (t_max - t_min)))

    def logsnr_schedule_cosine_shifted(self, t):
        return self.logsnr_schedule_cosine(t) + 2 * math.log(self.noise_d / self.image_d)

    def encode_text_prompt(self, text):
        if self.encoder_type == 'nn':
            return self.encoder(text).unsqueeze(1)
        return text

    # ---- helpers ------------------------------------------------------------------------------------------------
    def _dist(self):
        import torch.distributed as dist
        if getattr(self.config, "dcb_shard", None) == "timestep" and dist.is_available() and dist.is_initialized() \
                and dist.get_world_size() > 1:
            return dist, dist.get_rank(), dist.get_world_size()
        return None, 0, 1

    def _max_samples(self, H, W):
        v = getattr(self.config, "dcb_max_batch", None)
        if v:
            return int(v)
        # ~16 M pixels of denoiser batch per launch sequence (1024 samples at 128^2: the 8^2 / 16^2 layers then fill all 148
        # SMs; peak activation footprint ~40 GB of the 180 GB = ~2.5 KB per pixel of batch), bounded by half of what the
        # device can still give us (free + cached by torch's allocator).  Results do not depend on the chunk size.
        want = max(1, min(2048, (1 << 24) // (H * W)))
        if "DCB_MAX_BATCH" in os.environ:
            return int(os.environ["DCB_MAX_BATCH"])
        now = time.monotonic()     # cudaMemGetInfo costs ~1 ms of host time ahead of the first launch: ask once a second
        if self._mem_probe is None or now - self._mem_probe[0] > 1.0:
            try:
                free, _ = torch.cuda.mem_get_info()
                free += torch.cuda.memory_reserved() - torch.cuda.memory_allocated()
            except RuntimeError:
                free = None
            self._mem_probe = (now, free)
        free = self._mem_probe[1]
        if free is not None:
            want = max(1, min(want, int(0.5 * free / (2560.0 * H * W))))
        return want

    # ---- the hot path ---------------------------------------------------------------------------------------------
    @torch.no_grad()
    def classify(self, x, text=None, fast=False, t_all=None, eps_all=None):
        """``t_all`` [T,BS] / ``eps_all`` [T,BS,C,H,W] optionally inject pre-drawn noise (parity runs); by default
        t comes from the CPU generator exactly as the reference draws it and eps from the in-kernel Philox stream
        (``config.dcb_eps == 'torch'`` draws eps with torch.randn_like per step like the reference instead)."""
        cfg = self.config
        assert self.encoder_type is not None, "Encoder must be provided for classification."
        assert len(cfg.evaluation_per_stage) == cfg.n_stages, \
            "Number of evaluations per stage must match the number of stages."
        assert len(cfg.n_keep_per_stage) == cfg.n_stages, \
            "Number of classes to keep per stage must match the number of stages."
        assert cfg.n_keep_per_stage[-1] == 1, "Only one class should be selected at the end of the classification process."
        assert cfg.n_fast_classes <= cfg.classes and cfg.n_fast_classes >= 2, \
            "Number of fast classes must be less than or equal to the total number of classes. Must be at least 2."
        if not x.is_cuda:
            raise RuntimeError("dcb200.DiffusionClassifier.classify needs CUDA tensors; there is no CPU path")

        per_stage = [0] + list(cfg.evaluation_per_stage)
        BS, Cimg, H, W = x.shape
        dev = x.device
        T = per_stage[-1]
        errors = torch.full((BS, cfg.classes, T), torch.inf, device=dev)
        if fast:
            text = text.view(-1, 1).to(dev)
            classes = torch.arange(cfg.classes).repeat(BS, 1).to(dev)
            wrong = classes[(classes == text) == False].view(BS, -1)  # noqa: E712
            sel = torch.randint(0, wrong.shape[1], (BS, cfg.n_fast_classes - 1)).to(dev)
            classes = torch.cat((text, torch.gather(wrong, 1, sel)), dim=1)
        else:
            classes = torch.arange(cfg.classes).repeat(BS, 1).to(dev)

        net = self.ema.ema_model  # the reference scores with the EMA copy (:700)
        is_dit = isinstance(net, DiT)
        if not isinstance(net, (UNetCondition2D, DiT)):
            raise TypeError("backbone must be a dcb200.UNetCondition2D or dcb200.DiT")
        ctx = net.make_ctx(dev)
        pk = net.packed(ctx)
        xin = x.contiguous().float()
        v_param = self.pred_param == 'v'
        table = None
        if not is_dit:  # collapsed cross-attention bias per class, once per call (graphs keep their own copy)
            table = net.cross_attn_table(ctx, pk, E.cast(ctx, self.encoder.weight))
        use_graph = ctx.precision == "bf16" and getattr(cfg, "dcb_cuda_graph", None) is not False \
            and os.environ.get("DCB_CUDA_GRAPH", "1") != "0"
        patch = net.config.patch_size if is_dit else 1
        No = (patch * patch * Cimg) if is_dit else Cimg
        rows = (H // patch) * (W // patch)
        fused = (ctx.precision == "bf16") and rows % 128 == 0 and not getattr(cfg, "dcb_unfused_mse", False)
        eps_mode = getattr(cfg, "dcb_eps", None) or "philox"
        seed = int(torch.initial_seed() & 0x7FFFFFFFFFFFFFFF) + 0x9E3779B9 * self._eps_calls
        call_id = self._eps_calls
        self._eps_calls += 1
        dist, rank, world = self._dist()
        if dist is not None:    # one table, scored in slices: every rank uses rank 0's draws (see sync_from_rank0)
            seed = int(sync_from_rank0(dist, torch.tensor([seed], dtype=torch.int64), dev).item())
            if fast:
                classes = sync_from_rank0(dist, classes, dev)
        max_s = self._max_samples(H, W)
        self._graphs.cap = max(self._graphs.cap, 2 * cfg.n_stages + 6)

        for i in range(cfg.n_stages):
            start, end = per_stage[i], per_stage[i + 1]
            nj = end - start
            # t for every step of the stage from the CPU generator, in the reference's order (:688)
            if t_all is None:
                t_stage = torch.stack([torch.rand(BS) for _ in range(nj)])
            else:
                t_stage = t_all[start:end].detach().cpu().float()
            if dist is not None:
                t_stage = sync_from_rank0(dist, t_stage, dev).cpu()
            logsnr = self.schedule(t_stage).to(dev)                 # [nj, BS]
            alpha = torch.sqrt(torch.sigmoid(logsnr)).reshape(-1).contiguous()
            sigma = torch.sqrt(torch.sigmoid(-logsnr)).reshape(-1).contiguous()
            logsnr = logsnr.reshape(-1).float().contiguous()
            if eps_all is not None:
                eps_stage = eps_all[start:end].to(dev).float().reshape(nj * BS, Cimg, H, W).contiguous()
            elif eps_mode == "torch":
                eps_stage = torch.stack([torch.randn_like(xin) for _ in range(nj)]).reshape(nj * BS, Cimg, H, W)
            else:
                eps_stage = None
            nk = classes.shape[1]
            # U-Net: layers ahead of the first cross-attention are class-independent -> computed once per unit
            share = (not is_dit) and nk > 1 and getattr(cfg, "dcb_share_prefix", None) is not False \
                and os.environ.get("DCB_SHARE_PREFIX", "1") != "0"
            n_units = nj * BS
            lo, hi = shard_range(n_units, rank, world)
            chunk = max(1, max_s // nk)
            slab = torch.zeros(BS, cfg.classes, nj, device=dev) if dist is not None else None
            for u0 in range(lo, hi, chunk):
                U = min(chunk, hi - u0)
                units = torch.arange(u0, u0 + U, device=dev)
                img = (units % BS).to(torch.int32)
                jrel = units // BS
                cls = classes[img.long()]                              # [U, nk]
                cls32 = cls.reshape(-1).to(torch.int32).contiguous()
                pro = dict(patch=patch, eps=None if eps_stage is None else eps_stage[u0:u0 + U], seed=seed,
                           unit_id0=start * BS + u0, alpha=alpha[u0:u0 + U], sigma=sigma[u0:u0 + U], img=img,
                           want_target=True, v_param=v_param)
                if use_graph:
                    key = ("classify", pk.gen, is_dit, U, nk, Cimg, H, W, v_param, fused, share, str(dev))
                    hit = self._graphs.get(key)
                    gr = hit["obj"] if hit is not None else self._graphs.put(
                        key, _GraphedDenoiser(net, ctx, pk, is_dit, U, nk, Cimg, H, W, patch, v_param, fused, share),
                        id(net), pk.gen)
                    gr.set_table(table, call_id)
                    E.prologue(ctx, 1 if is_dit else 0, xin, U, 1 if share else nk, Cimg, H, W, pk.kpad_in,
                               a_out=gr.a_in, target_out=gr.target, **pro)
                    gr.logsnr.copy_(logsnr[u0:u0 + U])
                    gr.cls.copy_(cls32)
                    if v_param:
                        gr.scale.copy_(alpha[u0:u0 + U].repeat_interleave(nk))
                    gr.launch()
                    err = gr.err
                else:
                    a_in, target = E.prologue(ctx, 1 if is_dit else 0, xin, U, 1 if share else nk, Cimg, H, W,
                                              pk.kpad_in, **pro)
                    err = torch.empty(U * nk, device=dev, dtype=torch.float32)
                    mse = dict(target=target, div=nk, ld=No, err=err, fused=fused,
                               scale=alpha[u0:u0 + U].repeat_interleave(nk).contiguous() if v_param else None)
                    if is_dit:
                        net.run(ctx, pk, a_in, logsnr[u0:u0 + U], U, nk, cls32, mse=mse)
                    else:
                        net.run(ctx, pk, a_in, logsnr[u0:u0 + U], U, nk, H, W, table, xattn_idx=cls32, mse=mse,
                                share_prefix=share)
                b_idx = img.long().repeat_interleave(nk)
                j_idx = jrel.repeat_interleave(nk)
                if slab is None:
                    errors[b_idx, cls.reshape(-1), start + j_idx] = err      # reference :713-714
                else:
                    slab[b_idx, cls.reshape(-1), j_idx] = err
            if slab is not None:
                combine_stage_errors(errors, slab, classes, start, end, dist)  # one collective per stage
            num_keep = cfg.n_keep_per_stage[i]
            end_of_stage_errors = errors[:, :, :end].mean(dim=2)             # reference :719
            _, keep_indices = torch.topk(end_of_stage_errors, num_keep, dim=1, largest=False)
            classes = keep_indices
        assert classes.shape[1] == 1, "Only one class should be selected at the end of the classification process."
        self.last_errors = errors
        return classes[:, 0]

    # ---- helpers shared by sample / loss -----------------------------------------------------------------------
    def clip(self, x):
        return torch.clamp(x, -1, 1)

    def diffuse(self, x, alpha_t, sigma_t):
        """reference API (:100-117): returns (z_t, eps).  Used by callers outside the launch sequences (the hot paths
        fuse q_sample into dcb_prologue)."""
        eps_t = torch.randn_like(x)
        return alpha_t * x + sigma_t * eps_t, eps_t

    def _denoiser_setup(self, net, dev):
        is_dit = isinstance(net, DiT)
        if not isinstance(net, (UNetCondition2D, DiT)):
            raise TypeError("backbone must be a dcb200.UNetCondition2D or dcb200.DiT")
        ctx = net.make_ctx(dev)
        pk = net.packed(ctx)
        table = None if is_dit else net.cross_attn_table(ctx, pk, E.cast(ctx, self.encoder.weight))
        return is_dit, ctx, pk, table

    def _sampler_coefs(self, from_t, dev, call=0):
        """[sampling_steps + 1, 8] fp32 rows {c, a_t, a_s, s_t, s_s, sqrt(var), w, 0} and the logsnr fed to the denoiser per
        evaluation, computed exactly as ddpm_sampler_step does (:189-205).  The last row is the reference's separate
        "final step" (:268-288), which re-evaluates at steps[-2]."""
        n = int(self.config.sampling_steps)
        steps = torch.linspace(from_t, 0.0, n + 1)
        lt = torch.cat([self.schedule(steps[:-1]), self.schedule(steps[-2:-1])]).float()
        ls = torch.cat([self.schedule(steps[1:]), self.schedule(steps[-1:])]).float()
        c = -torch.special.expm1(lt - ls)
        a_t, a_s = torch.sqrt(torch.sigmoid(lt)), torch.sqrt(torch.sigmoid(ls))
        s_t, s_s = torch.sqrt(torch.sigmoid(-lt)), torch.sqrt(torch.sigmoid(-ls))
        sd = torch.sqrt((s_s ** 2) * c)
        w = torch.full_like(c, float(self.cfg_w))
        # Philox unit base of every evaluation as int bits: distinct per step and per sample() call under one seed
        step = (torch.arange(n + 1, dtype=torch.int64) + call * (n + 1)).remainder(1 << 30).to(torch.int32).view(torch.float32)
        return torch.stack([c, a_t, a_s, s_t, s_s, sd, w, step], 1).contiguous().to(dev), lt.to(dev)

    # ---- next row f2: DDPM ancestral sampler with classifier-free guidance (:175-293) ------------------------------
    @torch.no_grad()
    def ddpm_sampler_step(self, z_t, pred, u_pred, logsnr_t, logsnr_s):
        """reference signature (:176-207): (mu, variance).  mu comes from dcb_ddpm_step (noise-free, unclipped)."""
        dev = z_t.device
        lt, ls = logsnr_t.reshape(-1)[:1].float().cpu(), logsnr_s.reshape(-1)[:1].float().cpu()
        c = -torch.special.expm1(lt - ls)
        a_t, a_s = torch.sqrt(torch.sigmoid(lt)), torch.sqrt(torch.sigmoid(ls))
        s_t, s_s = torch.sqrt(torch.sigmoid(-lt)), torch.sqrt(torch.sigmoid(-ls))
        var = (s_s ** 2) * c
        coef = torch.cat([c, a_t, a_s, s_t, s_s, torch.sqrt(var), torch.tensor([float(self.cfg_w)]),
                          torch.zeros(1)]).float().to(dev)
        B, C, H, W = z_t.shape
        pair = torch.stack([pred, u_pred], 1).reshape(2 * B, C, H * W).transpose(1, 2).contiguous().float()  # NHWC rows
        ctx = E.Ctx(device=dev, precision="fp32")
        mu = E.ddpm_step(ctx, z_t.contiguous().float(), pair, 2, 0, coef, self.pred_param == 'v', False,
                         noise=torch.zeros_like(z_t, dtype=torch.float32))
        return mu, var.to(dev)

    @torch.no_grad()
    def sample(self, x, text=None, from_t=1, z_init=None, noise_all=None):
        """Same contract as the reference's ``sample(x, text, from_t)``.  Per step ONE denoiser launch sequence scores
        the conditional and the unconditional (null-token) branch as the two "classes" of every image -- the same batch
        folding (and, for the U-Net, the same shared class-independent prefix) as ``classify`` -- and dcb_ddpm_step fuses
        guidance, x-prediction, clipping, the posterior mean and the z_s draw.  With cfg_w == 0 the unconditional
        branch has weight exactly 0 and is not evaluated.  ``z_init`` / ``noise_all`` [steps,B,C,H,W] inject pre-drawn
        noise for parity runs (default: z_T as the reference draws it, per-step noise from the in-kernel Philox)."""
        if not x.is_cuda:
            raise RuntimeError("dcb200.DiffusionClassifier.sample needs CUDA tensors; there is no CPU path")
        if text is None or self.encoder_type is None:
            raise NotImplementedError("the reference's denoisers are class-conditional: sample() needs `text` labels")
        dev = x.device
        B, Cimg, H, W = x.shape
        net = self.ema.ema_model
        is_dit, ctx, pk, table = self._denoiser_setup(net, dev)
        if z_init is not None:
            z = z_init.to(dev).float().contiguous()
        elif from_t == 1:
            z = torch.randn(x.shape).to(dev)                              # :221, CPU generator like the reference
        else:
            logsnr = self.schedule(torch.ones(B) * from_t).to(dev)
            z, _ = self.diffuse(x.float(), torch.sqrt(torch.sigmoid(logsnr)).view(-1, 1, 1, 1),
                                torch.sqrt(torch.sigmoid(-logsnr)).view(-1, 1, 1, 1))
            z = z.contiguous()
        call = self._eps_calls
        self._eps_calls += 1
        coef, lt = self._sampler_coefs(from_t, dev, call)
        n = int(self.config.sampling_steps)
        guided = float(self.cfg_w) != 0.0
        rep = 2 if guided else 1
        text = text.to(dev).reshape(-1)
        cls = torch.stack([text, torch.full_like(text, self.null_token)], 1) if guided else text.reshape(-1, 1)
        cls32 = cls.reshape(-1).to(torch.int32).contiguous()
        share = (not is_dit) and guided
        patch = net.config.patch_size if is_dit else 1
        v_param = self.pred_param == 'v'
        seed = int(torch.initial_seed() & 0x7FFFFFFFFFFFFFFF)
        use_graph = ctx.precision == "bf16" and noise_all is None and getattr(self.config, "dcb_cuda_graph", None) is not False \
            and os.environ.get("DCB_CUDA_GRAPH", "1") != "0"
        mode = 1 if is_dit else 0
        rows = (H // patch) * (W // patch)
        if use_graph:   # persistent buffers: a replay reads / writes fixed addresses; z is updated in place
            key = ("sample", pk.gen, is_dit, B, rep, Cimg, H, W, v_param, share, str(dev))
            hit = self._graphs.get(key)
            st = hit["obj"] if hit is not None else self._graphs.put(key, dict(
                z=torch.empty(B, Cimg, H, W, device=dev), t=torch.empty(B, device=dev), coef=torch.empty(8, device=dev),
                a_in=ctx.empty((B if share else B * rep) * rows, pk.kpad_in), table=None, graph=None, warm=0, seed=0,
                cls=torch.empty(B * rep, device=dev, dtype=torch.int32),
                pk=pk, net=net), id(net), pk.gen)      # the graph's kernels point into pk's weights: keep them alive
            st["z"].copy_(z)
            if table is not None:
                if st["table"] is None or st["table"].shape != table.shape:
                    st["table"] = torch.empty_like(table)
                st["table"].copy_(table)
            st["cls"].copy_(cls32)

            def body(final):
                E.prologue(ctx, mode, st["z"], B, 1 if share else rep, Cimg, H, W, pk.kpad_in, patch=patch, a_out=st["a_in"])
                if is_dit:
                    pred = net.run(ctx, pk, st["a_in"], st["t"], B, rep, st["cls"])
                else:
                    pred = net.run(ctx, pk, st["a_in"], st["t"], B, rep, H, W, st["table"], xattn_idx=st["cls"],
                                   share_prefix=share)
                E.ddpm_step(ctx, st["z"], pred, rep, patch if is_dit else 0, st["coef"], v_param, final, seed=st["seed"],
                            out=st["z"])
            if st["graph"] is not None and st["seed"] != seed:
                st["graph"] = None      # the Philox seed is a kernel argument baked into the graph: capture again
            st["seed"] = seed
            for i in range(n + 1):
                st["t"].copy_(lt[i].expand(B))
                st["coef"].copy_(coef[i])
                if i == n:
                    body(True)                                   # the final (noise-free, clipped) step runs eagerly
                elif st["graph"] is not None:
                    st["graph"].replay()
                    L.lib().dcb_note_graph_replay(st["n_kernels"])
                elif st["warm"] < 1:
                    body(False)
                    st["warm"] += 1
                else:
                    torch.cuda.synchronize()
                    g = torch.cuda.CUDAGraph()
                    if _GraphedDenoiser._pool is None:
                        _GraphedDenoiser._pool = torch.cuda.graph_pool_handle()
                    n0 = L.launch_count()
                    with torch.cuda.graph(g, pool=_GraphedDenoiser._pool):
                        body(False)
                    st["n_kernels"] = L.launch_count() - n0
                    L.lib().dcb_note_graph_replay(-st["n_kernels"])
                    st["graph"] = g
                    g.replay()
                    L.lib().dcb_note_graph_replay(st["n_kernels"])
            return st["z"].clone()
        for i in range(n + 1):
            a_in, _ = E.prologue(ctx, mode, z, B, 1 if share else rep, Cimg, H, W, pk.kpad_in, patch=patch)
            t = lt[i].expand(B).contiguous()
            if is_dit:
                pred = net.run(ctx, pk, a_in, t, B, rep, cls32)
            else:
                pred = net.run(ctx, pk, a_in, t, B, rep, H, W, table, xattn_idx=cls32, share_prefix=share)
            final = i == n
            noise = None if (final or noise_all is None) else noise_all[i].to(dev).float().contiguous()
            z = E.ddpm_step(ctx, z, pred, rep, patch if is_dit else 0, coef[i], v_param, final, noise=noise, seed=seed)
        return z

    # ---- next row f4: training loss, forward only (:295-344) --------------------------------------------------------
    @torch.no_grad()
    def loss(self, x, text=None, t=None, eps=None):
        """min-SNR weighted eps-MSE of the ONLINE model (forward only: no autograd graph is built -- training's backward is
        outside this library).  q_sample is the fused prologue, the squared error the fused epilogue of the last GEMM."""
        if not x.is_cuda:
            raise RuntimeError("dcb200.DiffusionClassifier.loss needs CUDA tensors; there is no CPU path")
        if text is None or self.encoder_type is None:
            raise NotImplementedError("the reference's denoisers are class-conditional: loss() needs `text` labels")
        dev = x.device
        B, Cimg, H, W = x.shape
        t = torch.rand(B) if t is None else t.detach().cpu().float()
        net = self.model
        is_dit, ctx, pk, table = self._denoiser_setup(net, dev)
        logsnr = self.schedule(t).to(dev).float()
        alpha = torch.sqrt(torch.sigmoid(logsnr)).contiguous()
        sigma = torch.sqrt(torch.sigmoid(-logsnr)).contiguous()
        patch = net.config.patch_size if is_dit else 1
        No = (patch * patch * Cimg) if is_dit else Cimg
        rows = (H // patch) * (W // patch)
        v_param = self.pred_param == 'v'
        seed = int(torch.initial_seed() & 0x7FFFFFFFFFFFFFFF) + 0x9E3779B9 * self._eps_calls
        self._eps_calls += 1
        fused = (ctx.precision == "bf16") and rows % 128 == 0
        cls32 = text.to(dev).reshape(-1).to(torch.int32).contiguous()
        pro = dict(patch=patch, eps=None if eps is None else eps.to(dev).float().contiguous(), seed=seed, alpha=alpha,
                   sigma=sigma, want_target=True, v_param=v_param)
        use_graph = ctx.precision == "bf16" and getattr(self.config, "dcb_cuda_graph", None) is not False \
            and os.environ.get("DCB_CUDA_GRAPH", "1") != "0"
        if use_graph:      # same captured launch sequence as classify's, with one "class" (the label) per image
            key = ("loss", pk.gen, is_dit, B, Cimg, H, W, v_param, fused, str(dev))
            hit = self._graphs.get(key)
            gr = hit["obj"] if hit is not None else self._graphs.put(
                key, _GraphedDenoiser(net, ctx, pk, is_dit, B, 1, Cimg, H, W, patch, v_param, fused, False), id(net), pk.gen)
            gr.set_table(table, self._eps_calls)
            E.prologue(ctx, 1 if is_dit else 0, x.contiguous().float(), B, 1, Cimg, H, W, pk.kpad_in, a_out=gr.a_in,
                       target_out=gr.target, **pro)
            gr.logsnr.copy_(logsnr)
            gr.cls.copy_(cls32)
            if v_param:
                gr.scale.copy_(alpha)
            gr.launch()
            err = gr.err
        else:
            a_in, target = E.prologue(ctx, 1 if is_dit else 0, x.contiguous().float(), B, 1, Cimg, H, W, pk.kpad_in, **pro)
            err = torch.empty(B, device=dev, dtype=torch.float32)
            mse = dict(target=target, div=1, ld=No, err=err, fused=fused, scale=alpha if v_param else None)
            if is_dit:
                net.run(ctx, pk, a_in, logsnr.contiguous(), B, 1, cls32, mse=mse)
            else:
                net.run(ctx, pk, a_in, logsnr.contiguous(), B, 1, H, W, table, xattn_idx=cls32, mse=mse)
        snr = torch.exp(logsnr).clamp_(max=5)
        weight = 1 / (1 + snr) if v_param else 1 / snr
        return (weight * err).sum() / (B * Cimg * H * W)

    # ---- next row f1: callers of the hot paths (diffusion_classifier.py:532-655) -----------------------------------
    @staticmethod
    def _prefetched(loader, dev, stop_idx=None):
        """batches of ``loader`` with every tensor on ``dev``: batch k+1 is copied (persistent pinned staging buffers, side
        stream) while batch k is being scored, so the H2D never sits between two launch sequences and no pinned memory is
        allocated per batch."""
        side = torch.cuda.Stream(dev)
        it = iter(loader)
        staging = [{}, {}]          # two generations of pinned buffers per batch key; reused when shape / dtype match
        copied = [None, None]       # event of the last H2D that read generation g

        def load(gen):
            try:
                b = next(it)
            except StopIteration:
                return None
            if copied[gen] is not None:
                copied[gen].synchronize()            # the copy out of these staging buffers (two batches ago) is done
            out = {}
            with torch.cuda.stream(side):
                for k, v in b.items():
                    if torch.is_tensor(v) and not v.is_cuda:
                        if not v.is_pinned():
                            buf = staging[gen].get(k)
                            if buf is None or buf.shape != v.shape or buf.dtype != v.dtype:
                                buf = staging[gen][k] = torch.empty(v.shape, dtype=v.dtype, pin_memory=True)
                            buf.copy_(v)
                            v = buf
                        v = v.to(dev, non_blocking=True)
                    out[k] = v
            ev = torch.cuda.Event()
            ev.record(side)
            copied[gen] = ev
            return out, ev

        idx, nxt = 0, load(0)
        while nxt is not None:
            cur, ev = nxt
            last = stop_idx is not None and idx == stop_idx
            nxt = None if last else load((idx + 1) & 1)
            torch.cuda.current_stream(dev).wait_event(ev)
            for v in cur.values():
                if torch.is_tensor(v) and v.is_cuda:
                    v.record_stream(torch.cuda.current_stream(dev))
            yield cur
            idx += 1

    @torch.no_grad()
    def evaluate(self, val_dataloader, stop_idx=None, metrics=None, classification=False, from_t=1):
        """reference :532-578.  Differences in issue order only: the next batch's H2D overlaps the current batch's
        launch sequences, and nothing here forces a host sync (use dcb200.metrics for counters that stay on the GPU)."""
        val_samples, batches = [], []
        dev = next(self.ema.ema_model.parameters()).device
        for batch in self._prefetched(val_dataloader, dev, stop_idx):
            x = batch["images"]
            p = batch["prompt"] if "prompt" in batch.keys() else None
            if classification:
                sample = self.classify(x, p, fast=bool(self.config.fast_classification))
            else:
                sample = self.sample(x, p, from_t)
            if metrics is not None:
                for metric in metrics:
                    metric.update((sample, batch))
            val_samples.append(sample)
            batches.append(batch)
        return val_samples, batches, metrics

    # ---- next row f3: accelerate-layout checkpoints (diffusion_classifier.py:727-805) --------------------------------
    def _ckpt_modules(self):
        mods = [self.model, self.ema]                       # accelerator.prepare order (:383-388 / :617-621)
        if self.encoder_type == 'nn':
            mods.append(self.encoder)
        return mods

    def save_checkpoint(self, checkpoint_dir, epoch=0, best_metric=None, experiment_key=None):
        """writes what ``accelerator.save_state`` + the reference's experiment_state.pth write for the modules on the
        inference path: model.safetensors, model_1.safetensors (EMA), model_2.safetensors (encoder)."""
        from safetensors.torch import save_file
        os.makedirs(checkpoint_dir, exist_ok=True)
        for i, m in enumerate(self._ckpt_modules()):
            sd = {k: v.detach().cpu().contiguous().clone() for k, v in m.state_dict().items()}
            save_file(sd, os.path.join(checkpoint_dir, "model.safetensors" if i == 0 else f"model_{i}.safetensors"))
        torch.save({"epoch": epoch, "best_metric": best_metric, "experiment_key": experiment_key},
                   os.path.join(checkpoint_dir, "experiment_state.pth"))

    def load_checkpoint(self, checkpoint_path, accelerator=None):
        """reference :769-805 without accelerate: loads the state ``accelerator.save_state`` wrote (safetensors, or the
        older pytorch_model*.bin) into model / ema / encoder -- diffusers state_dict keys, SURVEY Appendix B -- and
        returns (epoch, best_metric, experiment_key).  Packed kernel weights are rebuilt lazily (parameter versions)."""
        for i, m in enumerate(self._ckpt_modules()):
            stem = "model" if i == 0 else f"model_{i}"
            f_st = os.path.join(checkpoint_path, stem + ".safetensors")
            f_bin = os.path.join(checkpoint_path, ("pytorch_model" if i == 0 else f"pytorch_model_{i}") + ".bin")
            if os.path.exists(f_st):
                from safetensors.torch import load_file
                sd = load_file(f_st)
            elif os.path.exists(f_bin):
                sd = torch.load(f_bin, map_location="cpu", weights_only=True)
            else:
                raise FileNotFoundError(f"no {stem}.safetensors / .bin under {checkpoint_path}")
            m.load_state_dict(sd)
        epoch, best_metric, experiment_key = 0, None, None
        state = os.path.join(checkpoint_path, "experiment_state.pth")
        if os.path.exists(state):
            ck = torch.load(state, map_location="cpu", weights_only=False)
            epoch, best_metric, experiment_key = ck["epoch"], ck.get("best_metric"), ck.get("experiment_key")
        return epoch, best_metric, experiment_key

    @torch.no_grad()
    def inference(self, optimizer=None, train_dataloader=None, val_dataloader=None, lr_scheduler=None, metrics=None,
                  plot_function=None, classification=False, from_t=1, checkpoint_folder="checkpoints"):
        """reference :580-655: load the most recent checkpoint, evaluate, sync the metrics, plot samples.  One process
        per GPU (torchrun) replaces accelerate: with ``config.dcb_shard == 'timestep'`` every rank sees every batch and
        ``classify`` shards the (image x timestep) units; otherwise ranks take batches round-robin (replicas)."""
        import torch.distributed as dist
        inference_image_path = os.path.join(self.config.experiment_path, "inference_images/")
        os.makedirs(inference_image_path, exist_ok=True)
        dev = torch.device("cuda", torch.cuda.current_device())
        self.to(dev)
        checkpoint_path = os.path.join(self.config.experiment_path, checkpoint_folder)
        self.load_checkpoint(checkpoint_path)
        if metrics is not None:
            for metric in metrics:
                metric.set_device(dev)
        self.model.eval()
        multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        replicas = multi and getattr(self.config, "dcb_shard", None) != "timestep"
        loader = val_dataloader
        if replicas:
            rank, world = dist.get_rank(), dist.get_world_size()
            loader = (b for i, b in enumerate(val_dataloader) if i % world == rank)
        val_samples, batches, metrics = self.evaluate(loader, metrics=metrics, stop_idx=self.config.evaluation_batches,
                                                      classification=classification, from_t=from_t)
        metric_output = []
        if metrics is not None:
            for metric in metrics:
                if replicas:
                    metric.sync_across_processes(None)
                metric_output.append(metric.get_output())
        if plot_function is not None and not classification:
            plot_function(output_dir=inference_image_path, batches=batches, samples=val_samples, epoch=0,
                          process_idx=dist.get_rank() if multi else 0)
        return (metric_output, val_samples, batches) if metrics is not None else (val_samples, batches)
