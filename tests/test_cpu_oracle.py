"""CPU suite (no GPU): pins the oracle against the reference's verbatim code / committed golden vectors / closed-form
KATs, and checks host logic + the C-ABI surface.  Run: python -m pytest tests -q -m "not gpu"."""
import math
import os
import re

import numpy as np
import pytest
import torch

from helpers import CIFAR_UNET, SMALL_UNET, TINY_DIT, TINY_UNET, UNET128, Cfg, base_cfg

GOLD = os.path.join(os.path.dirname(__file__), "golden")
ROOT = os.path.dirname(os.path.dirname(__file__))


def gold(name):
    return np.load(os.path.join(GOLD, name))


def checksum(m):
    return float(sum(p.detach().double().abs().sum() for p in m.parameters()))


# ---- schedule (diffusion_classifier.py:119-161) ------------------------------------------------------------
def test_schedule_kats_appendix_c():
    from oracle import loop
    t = torch.tensor([0, .25, .5, .75, 1.0])
    a = loop.logsnr_schedule_cosine(t, 32, 32)
    assert torch.allclose(a, torch.tensor([15.0, 1.7611834, 1.19e-07, -1.7611831, -14.999989]), atol=2e-5)
    b = loop.logsnr_schedule_cosine(t, 64, 256)
    assert torch.allclose(b, torch.tensor([13.613706, 1.7584484, -0.0016591818, -1.7631383, -16.386442]), atol=2e-5)
    c = loop.logsnr_schedule_cosine_shifted(t, 64, 256)
    assert torch.allclose(c, torch.tensor([10.841117, -1.0141404, -2.7742479, -4.5357270, -19.159031]), atol=2e-5)
    lam = loop.logsnr_schedule_cosine(torch.tensor([0.5]), 7, 7)
    assert abs(float(torch.sigmoid(lam)) - 0.5) < 1e-6  # alpha = sigma = sqrt(.5)


@pytest.mark.parametrize("tag,sched,nd,im", [("cos_32_32", "cosine", 32, 32), ("cos_64_256", "cosine", 64, 256),
                                              ("shift_64_256", "shifted_cosine", 64, 256),
                                              ("shift_32_128", "shifted_cosine", 32, 128)])
def test_schedule_matches_reference_golden(tag, sched, nd, im):
    """golden rows were produced by the reference's own DiffusionClassifier.schedule (oracle/make_golden.py)."""
    from oracle import loop
    import dcb200
    g = gold("schedule_kat.npz")
    t = torch.from_numpy(g["t"])
    cfg = base_cfg(schedule=sched, noise_d=nd, image_size=im)
    assert np.array_equal(loop.schedule_fn(cfg)(t).numpy(), g[tag])
    stub = torch.nn.Linear(1, 1)
    stub.config = type("c", (), {"encoder_hid_dim": 4})()
    dc = dcb200.DiffusionClassifier(stub, cfg)
    assert np.array_equal(dc.schedule(t).numpy(), g[tag])  # product host logic, bit-exact


# ---- denoiser restatement: structure ---------------------------------------------------------------------------
@pytest.mark.parametrize("arch,millions", [(CIFAR_UNET, 98.391939), (UNET128, 275.819523)])
def test_unet_param_counts_match_survey(arch, millions):
    from oracle.diffusers_restated import UNet2DConditionModel
    with torch.device("meta"):
        m = UNet2DConditionModel(**arch)
    assert abs(sum(p.numel() for p in m.parameters()) / 1e6 - millions) < 1e-6


def test_dit_b4_param_count():
    from oracle.diffusers_restated import DiTTransformer2DModel
    m = DiTTransformer2DModel(num_attention_heads=12, attention_head_dim=64, in_channels=3, out_channels=3,
                              num_layers=12, sample_size=256, patch_size=4)
    assert sum(p.numel() for p in m.parameters()) == 147476784


@pytest.mark.parametrize("kind,arch", [("unet", SMALL_UNET), ("unet", CIFAR_UNET), ("dit", TINY_DIT)])
def test_state_dict_schema_product_equals_oracle(kind, arch):
    from oracle import diffusers_restated as dr
    import dcb200
    o = (dr.UNet2DConditionModel if kind == "unet" else dr.DiTTransformer2DModel)(**arch)
    p = (dcb200.UNetCondition2D if kind == "unet" else dcb200.DiT)(**arch)
    so, sp = o.state_dict(), p.state_dict()
    assert set(so) == set(sp)
    assert all(so[k].shape == sp[k].shape for k in so)
    if kind == "unet":  # SURVEY Appendix B spot checks
        for k in ("conv_in.weight", "time_embedding.linear_1.weight", "encoder_hid_proj.bias",
                  "mid_block.attentions.0.transformer_blocks.0.attn2.to_v.weight",
                  "mid_block.attentions.0.transformer_blocks.0.ff.net.0.proj.weight",
                  "down_blocks.0.downsamplers.0.conv.weight", "up_blocks.0.upsamplers.0.conv.bias", "conv_norm_out.weight"):
            assert k in sp, k
        assert "mid_block.attentions.0.transformer_blocks.0.attn1.to_q.bias" not in sp
    else:
        for k in ("pos_embed.proj.weight", "transformer_blocks.0.norm1.emb.timestep_embedder.linear_1.weight",
                  "transformer_blocks.1.norm1.emb.class_embedder.embedding_table.weight",
                  "transformer_blocks.0.norm1.linear.bias", "transformer_blocks.0.attn1.to_q.bias", "proj_out_2.weight"):
            assert k in sp, k
        assert "pos_embed.pos_embed" not in sp
        assert torch.equal(o.pos_embed.pos_embed, p.pos_embed.pos_embed)


def test_timestep_embedding_kat():
    from oracle.diffusers_restated import get_timestep_embedding
    e = get_timestep_embedding(torch.zeros(1), 128, True, 0)
    assert torch.equal(e[0, :64], torch.ones(64)) and torch.equal(e[0, 64:], torch.zeros(64))
    e = get_timestep_embedding(torch.tensor([2.0]), 8, True, 0)
    w = torch.exp(-math.log(10000) * torch.arange(4) / 4)
    assert torch.allclose(e[0], torch.cat([torch.cos(2 * w), torch.sin(2 * w)]))


def test_single_token_cross_attention_is_a_bias():
    """softmax over one key == 1  =>  attn2(h, ctx) == to_out(to_v(ctx)) for any h (SURVEY finding 3)."""
    from oracle.diffusers_restated import Attention
    torch.manual_seed(0)
    a = Attention(64, 8, 8, cross_attention_dim=32)
    h, ctx = torch.randn(2, 10, 64), torch.randn(2, 1, 32)
    out = a(h, ctx)
    bias = a.to_out[0](a.to_v(ctx))
    assert torch.allclose(out, bias.expand_as(out), atol=1e-6)


# ---- Haar ------------------------------------------------------------------------------------------------------
def test_haar_oracle_kats_and_golden():
    from oracle import haar
    lo, hi = haar.dwt(np.array([1.0, 2.0, 3.0, 4.0]))
    assert np.allclose(lo, [2.12132034, 4.94974747]) and np.allclose(hi, [-0.70710678, -0.70710678])  # pywt docs
    cA, (cH, cV, cD) = haar.dwt2(np.array([[1.0, 2.0], [3.0, 4.0]]))
    assert np.allclose([cA, cH, cV, cD], [[[5.0]], [[-2.0]], [[-1.0]], [[0.0]]])
    g = gold("haar_kat.npz")  # produced by the reference's utils/wavelet.py over this Haar
    assert np.allclose(haar.wavelet_dec_2_np(g["img"]), g["w"], atol=1e-7)
    assert np.allclose(haar.wavelet_enc_2_np(g["w"]), g["img"], atol=1e-6)
    assert np.allclose(g["back"], g["img"], atol=1e-6)


# ---- loop: restated oracle == golden produced by the reference's verbatim classify --------------------------------
def _build(kind, arch, cfg, seed, factor):
    from oracle import diffusers_restated as dr
    torch.manual_seed(seed)
    net = (dr.UNet2DConditionModel if kind == "unet" else dr.DiTTransformer2DModel)(**arch)
    enc = torch.nn.Embedding(cfg.classes + 1, arch["encoder_hid_dim"]) if kind == "unet" else None
    with torch.no_grad():
        if kind == "unet":
            enc.weight.mul_(factor)
        else:
            for b in net.transformer_blocks:
                b.norm1.emb.class_embedder.embedding_table.weight.mul_(factor)
    return net.eval(), enc


FIXTURES = {
    "classify_unet_tiny.npz": ("unet", TINY_UNET, dict(pred_param="eps", schedule="cosine", noise_d=16, image_size=16,
                                                         encoder_type="nn", classes=4, n_stages=2,
                                                         evaluation_per_stage=[2, 4], n_keep_per_stage=[2, 1])),
    "classify_unet_small_v.npz": ("unet", SMALL_UNET, dict(pred_param="v", schedule="shifted_cosine", noise_d=16,
                                                            image_size=32, encoder_type="nn", classes=2, n_stages=1,
                                                            evaluation_per_stage=[3], n_keep_per_stage=[1])),
    "classify_dit_tiny.npz": ("dit", TINY_DIT, dict(pred_param="v", schedule="shifted_cosine", noise_d=16, image_size=32,
                                                     encoder_type="DiT", classes=3, n_stages=2,
                                                     evaluation_per_stage=[2, 3], n_keep_per_stage=[2, 1])),
}


@pytest.mark.parametrize("name", list(FIXTURES))
def test_oracle_loop_reproduces_reference_golden(name):
    from oracle import loop
    kind, arch, kw = FIXTURES[name]
    g = gold(name)
    cfg = base_cfg(**kw)
    net, enc = _build(kind, arch, cfg, int(g["seed"]), float(g["factor"]))
    if abs(checksum(net) - float(g["checksum"])) > 1e-6 * float(g["checksum"]):
        pytest.skip("torch default-init stream differs from the build that wrote the fixture")

    class Den(torch.nn.Module):
        def forward(self, x, noise_labels, encoder_hidden_states):
            return net(x, noise_labels, encoder_hidden_states)[0]

    labels, errors = loop.classify_oracle(Den(), enc, cfg, torch.from_numpy(g["x"]), t_all=torch.from_numpy(g["t_all"]),
                                          eps_all=torch.from_numpy(g["eps_all"]), return_errors=True)
    assert labels.tolist() == g["labels"].tolist()
    final = errors.mean(dim=2).numpy()
    ref = g["stage_means"][-1]
    fin = np.isfinite(ref)
    assert np.array_equal(np.isfinite(final), fin)
    assert np.allclose(final[fin], ref[fin], rtol=1e-5)


def test_oracle_loop_equals_verbatim_reference_with_toy_backbone():
    """direct A/B against /root/reference's classify (imported verbatim), incl. 3-stage pruning and fast mode."""
    from oracle import loop
    from oracle.reference_loader import Config, injected_noise, load_reference, reference_available
    if not reference_available():
        pytest.skip("/root/reference not present (GPU box)")
    ref = load_reference()

    class Toy(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.randn(8, 3))
            self.config = type("c", (), {"encoder_hid_dim": 8})()

        def forward(self, x, noise_labels, encoder_hidden_states):
            s = torch.tanh(encoder_hidden_states[:, 0] @ self.w)  # [B,3]
            return x * s[:, :, None, None] + noise_labels.view(-1, 1, 1, 1) * 0.01

    for pred, sched, fast in (("eps", "cosine", False), ("v", "shifted_cosine", False), ("eps", "cosine", True)):
        torch.manual_seed(0)
        cfg = Config(pred_param=pred, schedule=sched, noise_d=16, image_size=8, cfg_w=0, ema_beta=0.99, ema_warmup=0,
                     ema_update_freq=1, encoder_type="nn", classes=6, n_stages=3, evaluation_per_stage=[2, 4, 7],
                     n_keep_per_stage=[4, 2, 1], n_fast_classes=3)
        dc = ref.DiffusionClassifier(Toy(), cfg).eval()
        x = torch.rand(5, 3, 8, 8) * 2 - 1
        text = torch.randint(0, 6, (5,))
        t_all, eps_all = torch.rand(7, 5), torch.randn(7, 5, 3, 8, 8)
        torch.manual_seed(123)
        with injected_noise(dc, t_all, eps_all) as st:
            y_ref = dc.classify(x, text, fast=fast)
        torch.manual_seed(123)
        y, errors = loop.classify_oracle(dc.ema, dc.encoder, cfg, x, text, fast=fast, t_all=t_all, eps_all=eps_all,
                                         return_errors=True)
        assert torch.equal(y, y_ref)
        assert torch.equal(errors.mean(dim=2), st["stage_means"][-1])


def test_reference_wrappers_forward_golden():
    """restated denoisers reproduce the outputs recorded through the reference's nets/unet.py / nets/dit.py wrappers."""
    from oracle import diffusers_restated as dr
    g = gold("unet_small_forward.npz")
    torch.manual_seed(11)
    u = dr.UNet2DConditionModel(**SMALL_UNET).eval()
    if abs(checksum(u) - float(g["checksum"])) < 1e-6 * float(g["checksum"]):
        with torch.no_grad():
            y = u(torch.from_numpy(g["x"]), torch.from_numpy(g["lam"]), torch.from_numpy(g["ehs"]))[0]
        assert np.allclose(y.numpy(), g["y"], atol=1e-5)
    g = gold("dit_tiny_forward.npz")
    torch.manual_seed(13)
    d = dr.DiTTransformer2DModel(**TINY_DIT).eval()
    if abs(checksum(d) - float(g["checksum"])) < 1e-6 * float(g["checksum"]):
        with torch.no_grad():
            y = d(torch.from_numpy(g["x"]), torch.from_numpy(g["lam"]), torch.from_numpy(g["lab"]))[0]
        assert np.allclose(y.numpy(), g["y"], atol=1e-5)


# ---- C ABI + host logic ---------------------------------------------------------------------------------------------
def test_c_abi_exports_every_declared_symbol():
    from dcb200 import _lib
    header = open(os.path.join(ROOT, "include", "dcb200.h")).read()
    declared = set(re.findall(r"\b(dcb_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    lib = _lib.lib()  # dlopen + bind every prototype (no compute without a GPU)
    for name in declared:
        assert hasattr(lib, name)
    assert lib.dcb_version() == 111 and lib.dcb_launch_count() == 0
    import ctypes
    assert ctypes.sizeof(_lib.Seg) == 48 == lib.dcb_struct_size(0)
    assert ctypes.sizeof(_lib.GemmDesc) == 808 == lib.dcb_struct_size(1)


def test_product_has_no_cpu_path():
    import dcb200
    torch.manual_seed(0)
    net = dcb200.UNetCondition2D(**TINY_UNET)
    with pytest.raises(RuntimeError, match="CUDA"):
        net(torch.zeros(1, 3, 16, 16), torch.zeros(1), encoder_hidden_states=torch.zeros(1, 1, 64))
    dc = dcb200.DiffusionClassifier(net, base_cfg())
    with pytest.raises(RuntimeError, match="CUDA"):
        dc.classify(torch.zeros(2, 3, 16, 16))
    with pytest.raises(AssertionError):  # reference's assertion on inconsistent stage config (:659-663)
        dcb200.DiffusionClassifier(net, base_cfg(n_stages=2)).classify(torch.zeros(2, 3, 16, 16))
    import importlib.util
    src = open(os.path.join(ROOT, "diffusion-classifier_b200", "dcb200", "classifier.py")).read()
    assert "oracle" not in src.replace("oracle/", "")  # the product never imports the checker


def test_ema_shim_contract():
    import dcb200
    m = torch.nn.Linear(4, 4)
    e = dcb200.EMA(m, beta=0.5, update_after_step=0, update_every=1)
    assert set(e.state_dict()) == {"initted", "step", "online_model.weight", "online_model.bias", "ema_model.weight",
                                   "ema_model.bias"}
    assert torch.equal(e(torch.ones(1, 4)), e.ema_model(torch.ones(1, 4)))


def test_ema_update_schedule():
    """ema_pytorch 0.7.7's update(): copies (without setting ``initted``) while step <= update_after_step, first
    post-warm-up call copies + sets initted, then lerps with decay = clamp(1 - (1 + epoch) ** (-2/3), min_value, beta),
    epoch = step_after_increment - update_after_step - 1; update_every gates everything; float buffers are averaged,
    integer buffers are not.  Hand-derived values (inv_gamma = 1, power = 2/3)."""
    import dcb200

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.zeros(3))
            self.register_buffer("fb", torch.zeros(2))
            self.register_buffer("ib", torch.zeros(2, dtype=torch.long))

    m = Net()
    e = dcb200.EMA(m, beta=0.9, update_after_step=2, update_every=1)

    def bump():
        with torch.no_grad():
            m.w.add_(1.0)
            m.fb.add_(2.0)
            m.ib.add_(1)

    for k in range(3):                      # steps 0, 1, 2: plain copies, still not initted
        bump()
        e.update()
        assert torch.equal(e.ema_model.w, m.w) and torch.equal(e.ema_model.fb, m.fb) and not bool(e.initted)
    bump()
    e.update()                              # step 3: first real update -> copy, initted, lerp of equal tensors
    assert bool(e.initted) and torch.equal(e.ema_model.w, m.w) and int(e.step) == 4
    assert abs(e.get_current_decay() - (1 - 2 ** (-2 / 3))) < 1e-12
    bump()
    e.update()                              # step 4: epoch = 5 - 2 - 1 = 2 -> decay = 1 - 3^(-2/3) = 0.5192499
    d = 1 - 3 ** (-2 / 3)
    assert torch.allclose(e.ema_model.w, m.w - 1.0 + (1 - d), atol=1e-6)
    assert torch.allclose(e.ema_model.fb, m.fb - 2.0 + 2 * (1 - d), atol=1e-6)
    assert torch.equal(e.ema_model.ib, torch.zeros(2, dtype=torch.long))      # integer buffers are never touched
    # decay is capped by beta and floored by min_value; update_every skips the steps in between
    e2 = dcb200.EMA(Net(), beta=0.4, update_after_step=0, update_every=3, min_value=0.39)
    for _ in range(4):
        e2.update()                         # steps 0 (copy), 1, 2 (skipped), 3 (init + lerp)
    assert bool(e2.initted) and int(e2.step) == 4
    assert 0.39 <= e2.get_current_decay() <= 0.4
    e2.step.fill_(1000)
    assert e2.get_current_decay() == 0.4


def _gloo_worker(rank, world, port, out):
    import torch.distributed as dist
    from dcb200.classifier import combine_stage_errors, shard_range
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    BS, classes, nj = 3, 5, 4
    alive = torch.tensor([[0, 3], [1, 2], [4, 0]])
    truth = torch.arange(BS * classes * nj, dtype=torch.float32).reshape(BS, classes, nj) + 0.5
    errors = torch.full((BS, classes, nj + 2), torch.inf)
    slab = torch.zeros(BS, classes, nj)
    lo, hi = shard_range(nj * BS, rank, world)
    for u in range(lo, hi):
        j, b = divmod(u, BS)
        for c in alive[b].tolist():
            slab[b, c, j] = truth[b, c, j]
    combine_stage_errors(errors, slab, alive, 2, 2 + nj, dist)
    if rank == 0:
        torch.save(errors, out)
    dist.destroy_process_group()


def test_timestep_shard_allreduce_gloo_world2(tmp_path):
    """N>1 host logic on CPU: 2 gloo ranks each own half of the (j,b) units; one all-reduce rebuilds the stage table."""
    import torch.multiprocessing as mp
    from dcb200.classifier import shard_range
    assert [shard_range(10, r, 4) for r in range(4)] == [(0, 2), (2, 5), (5, 7), (7, 10)]
    out = str(tmp_path / "e.pt")
    port = 29500 + os.getpid() % 2000
    mp.spawn(_gloo_worker, args=(2, port, out), nprocs=2, join=True)
    errors = torch.load(out)
    BS, classes, nj = 3, 5, 4
    alive = torch.tensor([[0, 3], [1, 2], [4, 0]])
    truth = torch.arange(BS * classes * nj, dtype=torch.float32).reshape(BS, classes, nj) + 0.5
    assert torch.isinf(errors[:, :, :2]).all()
    for b in range(BS):
        for c in range(classes):
            if c in alive[b].tolist():
                assert torch.equal(errors[b, c, 2:], truth[b, c])
            else:
                assert torch.isinf(errors[b, c, 2:]).all()


def _sync_worker(rank, world, port, out):
    import torch.distributed as dist
    from dcb200.classifier import sync_from_rank0
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    torch.manual_seed(100 + rank)                 # the usual "seed + rank" set-up: every rank draws differently
    BS, n_cls, n_fast = 4, 7, 3
    text = torch.tensor([1, 5, 0, 6]).view(-1, 1)
    classes = torch.arange(n_cls).repeat(BS, 1)   # fast-mode candidate draw, as classify() / reference :671-677
    wrong = classes[(classes == text) == False].view(BS, -1)  # noqa: E712
    sel = torch.randint(0, wrong.shape[1], (BS, n_fast - 1))
    mine = torch.cat((text, torch.gather(wrong, 1, sel)), dim=1)
    t_mine = torch.stack([torch.rand(BS) for _ in range(3)])
    seed_mine = torch.tensor([torch.initial_seed()], dtype=torch.int64)
    got = [sync_from_rank0(dist, v.clone(), torch.device("cpu")) for v in (mine, t_mine, seed_mine)]
    torch.save(dict(mine=(mine, t_mine, seed_mine), got=got), out + str(rank))
    dist.destroy_process_group()


def test_rank_consistency_sync_gloo_world2(tmp_path):
    """dcb_shard='timestep': ranks seeded differently must score ONE table -- rank 0's timesteps, fast-mode candidate
    classes and Philox seed reach every rank (classifier.sync_from_rank0); without it the all-reduced slab would mix
    class columns.  Also checks that the draws really differed before the sync (the test would be vacuous otherwise)."""
    import torch.multiprocessing as mp
    out = str(tmp_path / "s")
    port = 31500 + os.getpid() % 2000
    mp.spawn(_sync_worker, args=(2, port, out), nprocs=2, join=True)
    r0, r1 = torch.load(out + "0"), torch.load(out + "1")
    assert not torch.equal(r0["mine"][0], r1["mine"][0]) and not torch.equal(r0["mine"][1], r1["mine"][1])
    for a, b, m in zip(r0["got"], r1["got"], r0["mine"]):
        assert torch.equal(a, b) and torch.equal(a, m)


def test_graph_cache_drops_stale_pack_generations():
    """CUDA-graph cache: keyed on a monotonic pack generation, entries of a superseded generation of the same
    network go first, LRU otherwise; capacity is never exceeded."""
    from dcb200.classifier import _GraphCache
    c = _GraphCache()
    c.cap = 4
    for i in range(3):
        c.put(("k", 1, i), f"g1_{i}", net_id=7, gen=1)
    c.put(("other", 9, 0), "o", net_id=8, gen=9)
    assert len(c) == 4 and c.get(("k", 1, 0))["obj"] == "g1_0"
    c.put(("k", 2, 0), "g2_0", net_id=7, gen=2)          # repack of net 7: all gen-1 graphs of net 7 are dropped
    assert len(c) == 2 and c.get(("k", 1, 0)) is None and c.get(("other", 9, 0)) is not None
    for i in range(1, 6):
        c.put(("k", 2, i), f"g2_{i}", net_id=7, gen=2)
    assert len(c) == 4 and c.get(("other", 9, 0)) is None      # least recently used went first


def test_c_abi_argument_validation_returns_codes_not_crashes():
    """error behaviour of the boundary: bad arguments are rejected with DCB_EINVAL (-1) and a message in dcb_last_error()
    before anything is launched (so this runs without a GPU); nothing throws or exits across the ABI."""
    import ctypes as C
    from dcb200 import _lib
    lib = _lib.lib()

    def err():
        return lib.dcb_last_error().decode()

    d = _lib.GemmDesc()
    assert lib.dcb_gemm(None, None) == -1 and "null descriptor" in err()
    d.dtype, d.nseg = 7, 1
    assert lib.dcb_gemm(C.byref(d), None) == -1 and "dtype" in err()
    d.dtype, d.nseg = _lib.BF16, 0
    assert lib.dcb_gemm(C.byref(d), None) == -1 and "nseg" in err()
    d.nseg, d.NB, d.OH, d.OW, d.N = 1, 1, 1, 128, 64
    assert lib.dcb_gemm(C.byref(d), None) == -1 and "weights" in err()
    ok = C.c_int32(5)
    assert lib.dcb_gemm_gn_layout(C.byref(d), C.byref(ok)) == -1
    assert lib.dcb_ddpm_step(1, 1, 3, 0, 1, 0, 0, None, 0, 0, 1, 3, 8, 8, 1, None) == -1 and "rep" in err()
    assert lib.dcb_ddpm_step(None, 1, 2, 0, 1, 0, 0, None, 0, 0, 1, 3, 8, 8, 1, None) == -1 and "null" in err()
    assert lib.dcb_prologue(2, _lib.BF16, 1, None, 0, 0, None, None, None, 1, 1, 3, 8, 8, 1, 64, 1, 1, None, 0, None) == -1
    assert "mode" in err()
    assert lib.dcb_prologue(0, _lib.BF16, 1, None, 0, 0, None, None, None, 1, 1, 3, 8, 8, 1, 20, 1, 1, None, 0, None) == -1
    assert "kpad" in err()
    assert lib.dcb_groupnorm_fused(_lib.BF16, 1, 64, 1, None, 0, 1, 2, 64, 32, 1, 1, 1e-5, 1, 1, None) == -1   # 2 channels / group
    assert "channels per group" in err()
    assert lib.dcb_groupnorm_apply(_lib.BF16, 1, 60, None, 0, 2, 64, 32, 1, 1, 1, 1, 1e-5, 1, 1, None) == -1
    assert lib.dcb_layernorm(_lib.BF16, 1, 4, 2048, None, None, 1e-5, None, None, 0, 0, 1, None) == -1 and "layernorm" in err()
    assert lib.dcb_attention(_lib.BF16, 1, 1, 1, 192, 1, 128, 1, 48, 1.0, 1, 48, None) == -1 and "head dim" in err()
    assert lib.dcb_timestep_embed(_lib.BF16, 1, 1, 1, 7, 0.0, 10000.0, 1, None) == -1 and "even" in err()
    assert lib.dcb_eps_mse(_lib.BF16, 1, 1, None, 1, 0, 16, 1, 1, None) == -1 and "div" in err()
    assert lib.dcb_upsample2x(_lib.BF16, 1, 1, 4, 4, 12, 1, None) == -1
    assert lib.dcb_expand_samples(_lib.BF16, 1, 2, 1, 3, 1, None) == -1
    assert lib.dcb_pack_conv(_lib.BF16, 1, 8, 4, 3, 3, 20, 1, None) == -1 and "kpad" in err()
    assert lib.dcb_pack_geglu(_lib.BF16, 1, None, 100, 8, 1, None, None) == -1 and "multiple of 128" in err()
    assert lib.dcb_pack_upsample(7, 1, 8, 8, 1, None) == -1 and "dtype" in err()
    assert lib.dcb_pack_rows(_lib.BF16, 1, 4, 2, 8, 1, 8, 0, 0, None) == -1
    assert lib.dcb_launch_count() == 0          # nothing was launched


def test_params_version_sees_every_kind_of_weight_change():
    """engine.params_version (the key of the packed-weight and CUDA-graph caches, evaluated at the top of every classify()
    call) changes on an in-place update, on a replaced Parameter, on a replaced sub-module and on a moved storage -- and
    is stable otherwise."""
    import torch
    import torch.nn as nn
    from dcb200 import engine as E
    net = nn.Sequential(nn.Linear(4, 4), nn.Sequential(nn.Linear(4, 4), nn.LayerNorm(4)))
    v0 = E.params_version(net)
    assert v0 == E.params_version(net) and len(v0) == 2 * len(list(net.parameters()))
    with torch.no_grad():
        net[1][0].bias.add_(1.0)                       # optimizer step / load_state_dict / EMA copy: in place
    v1 = E.params_version(net)
    assert v1 != v0
    net[0].weight = nn.Parameter(net[0].weight.detach().clone())     # replaced Parameter (new storage)
    v2 = E.params_version(net)
    assert v2 != v1
    net[1][1] = nn.LayerNorm(4)                       # replaced sub-module
    v3 = E.params_version(net)
    assert v3 != v2
    net[0].weight.data = net[0].weight.data.clone()   # what Module.to() / .half() do: same Parameter, new storage
    assert E.params_version(net) != v3
