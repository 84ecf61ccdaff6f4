"""The CUDA-core kernels of ``csrc/fused_ops.cu`` and ``csrc/loss_ops.cu`` (fused optimizer, importance accumulation,
casts, label-smoothing CE, NHWC batch norm forward / backward, pooling, scalar window attention forward / backward,
triplet mining, KD-KL, BCE-distill) executed on the CPU under the SIMT emulator of ``tests/emu``.

These kernels are validated on the device by ``tests/test_gpu_kernels.py``; here their *source* runs in the CPU tier
through the same wrappers (``flpr_b200.ops.fused``) and entry points, so a change to a kernel or to its argument
marshalling is caught without a GPU. Every check calls the op twice - PyTorch reference path, then emulated kernel -
and compares values and gradients. (The tensor-core window-attention kernel in the same source file compiles against
aborting stubs and is never launched here.)"""
import shutil

import pytest
import torch
import torch.nn.functional as F


@pytest.fixture(scope="module")
def emu_paths():
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    from emu.build_emu import build
    try:
        return [build("fused_ops.cu"), build("loss_ops.cu")]
    except RuntimeError as ex:
        pytest.skip(f"emulator build unavailable: {ex}")


class both:
    """``with both(paths) as run: ref, got = run(lambda: op(...))``: the callable is evaluated on the reference path and
    then on the emulated kernels (fresh inputs each time: it must build them itself)."""

    def __init__(self, paths):
        self.paths = paths

    def __enter__(self):
        from flpr_b200.ops import fused as fops, native
        self.fops, self.native = fops, native

        def run(fn):
            fops.use_emulated_libraries(None)
            ref = fn()
            fops.use_emulated_libraries(self.paths)
            before = native.launches()
            try:
                got = fn()
            finally:
                fops.use_emulated_libraries(None)
            assert native.launches() > before, "the emulated run did not launch a kernel"
            return ref, got
        return run

    def __exit__(self, *exc):
        self.fops.use_emulated_libraries(None)
        return False


def close(a, b, rtol=1e-4, atol_frac=1e-5):
    a, b = a.float(), b.float()
    assert a.shape == b.shape
    atol = atol_frac * float(b.abs().max()) + 1e-7
    assert torch.allclose(a, b, rtol=rtol, atol=atol), float((a - b).abs().max())


def rand(*shape, seed=0, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale


# ------------------------------------------------------------------------------------------------------------ optimizer
@pytest.mark.parametrize("kind,penalty,l1", [("adam", False, False), ("adam", True, True), ("sgd", True, False),
                                             ("sgd", False, True)])
def test_fused_optimizer_step(emu_paths, kind, penalty, l1):
    n = 4 * 777

    def step():
        from flpr_b200.ops import fused as fops
        p, g = rand(n, seed=1), rand(n, seed=2, scale=0.1)
        m, v = rand(n, seed=3, scale=0.01), rand(n, seed=4, scale=0.01).abs()
        Q, R = (rand(n, seed=5).abs(), rand(n, seed=6)) if penalty else (None, None)
        G = rand(n, seed=7) if l1 else None
        shadow = torch.zeros(n, dtype=torch.bfloat16)
        stats = torch.zeros(4)
        for it in (1, 2):
            fops.fused_optimizer_step(kind, p, g, m, v if kind == "adam" else None, lr=1e-2, step=it, weight_decay=1e-3,
                                      momentum=0.9 if kind == "sgd" else 0.0, Q=Q, R=R, lam2=0.5 if penalty else 0.0,
                                      G=G, lam1=1e-2 if l1 else 0.0, atten=0.9 if l1 else 0.0, p_bf16=shadow,
                                      stats=stats)
        return p, m, v, shadow, stats
    with both(emu_paths) as run:
        ref, got = run(step)
    for a, b in zip(got[:3], ref[:3]):
        close(a, b)
    close(got[3], ref[0].to(torch.bfloat16), rtol=1e-2, atol_frac=1e-2)        # (the reference path does not refresh it)
    close(got[4][:2], ref[4][:2], rtol=1e-3, atol_frac=1e-4)


@pytest.mark.parametrize("kind", ["adam", "sgd"])
def test_trained_l1_anchor_kernel_matches_the_tensor_op_twin(emu_paths, kind):
    """FedSTIL's trained L1 anchor (reference quirk, ``fedstil.py:53-76,639-647``): ``fused_opt_kernel<true>`` under the
    emulator vs ``ArenaOptimizer._anchor_step`` (the fp32 tensor-op form the golden tests compare with the reference)
    over four steps - weights, anchor, both moment pairs, the reported L1 sum."""
    import torch.nn as nn

    def steps():
        from flpr_b200.runtime.arena import ArenaOptimizer, ParamArena
        torch.manual_seed(11)
        lin1, lin2 = nn.Linear(64, 96, bias=False), nn.Linear(96, 16)
        params = [("a.weight", lin1.weight), ("b.weight", lin2.weight), ("b.bias", lin2.bias)]
        arena = ParamArena(params, "cpu", shadow=True, first=lambda n: n == "a.weight")
        opt = ArenaOptimizer(kind, arena, lr=1e-3 if kind == "adam" else 0.05, weight_decay=1e-4, momentum=0.9)
        n = arena.prefix_numel
        torch.manual_seed(12)
        G = arena.master[:n] + 0.01 * torch.randn(n)
        opt.G, opt.lam1, opt.atten = G, 1e-2, 0.9
        opt.anchor = G.clone()
        opt.stats = torch.zeros(2)
        for step in range(4):
            torch.manual_seed(100 + step)
            arena.grad.copy_(torch.randn(arena.numel) * (torch.rand(arena.numel) > 0.3))     # exact zeros: sign(0) = 0
            opt.step()
        return (arena.master.clone(), opt.anchor.clone(), opt.m.clone(),
                opt.anchor_m.clone() if opt.anchor_m is not None else torch.zeros(1), opt.stats.clone(), G)
    with both(emu_paths) as run:
        ref, got = run(steps)
    close(got[0], ref[0], rtol=1e-4, atol_frac=1e-5)
    close(got[1], ref[1], rtol=1e-4, atol_frac=1e-5)
    close(got[2], ref[2], rtol=1e-3, atol_frac=1e-4)
    close(got[3], ref[3], rtol=1e-3, atol_frac=1e-4)
    close(got[4][1:], ref[4][1:], rtol=1e-3, atol_frac=1e-3)
    assert float((ref[1] - ref[5]).abs().max()) > 1e-5                      # the anchor did move


def test_importance_cast_and_compose(emu_paths):
    n = 4 * 501

    def ops():
        from flpr_b200.ops import fused as fops
        f1, f2, g = rand(n, seed=1).abs(), rand(n, seed=1).abs(), rand(n, seed=2)
        fops.importance_accumulate(f1, g, 0.25, "fisher")
        fops.importance_accumulate(f2, g, 0.25, "mas")
        c = fops.cast_bf16(g)
        theta, t16 = torch.zeros(n), torch.zeros(n, dtype=torch.bfloat16)
        fops.compose_adaptive(rand(n, seed=3), rand(n, seed=4), 0.9, theta, t16)
        return f1, f2, c, theta, t16
    with both(emu_paths) as run:
        ref, got = run(ops)
    close(got[0], ref[0])
    close(got[1], ref[1])
    assert torch.equal(got[2], ref[2])
    close(got[3], ref[3])
    close(got[4], ref[4], rtol=1e-2, atol_frac=1e-2)


# ------------------------------------------------------------------------------------------------------------ CE
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_label_smoothing_cross_entropy(emu_paths, dtype):
    def ce():
        from flpr_b200.ops import fused as fops
        logits = rand(24, 1000, seed=1, scale=3.0).to(dtype).requires_grad_(True)
        target = torch.randint(0, 1000, (24,), generator=torch.Generator().manual_seed(2))
        stats = torch.zeros(2)
        loss = fops.ce_label_smooth(logits, target, 0.1, stats)
        loss.backward()
        return loss.detach(), logits.grad, stats
    with both(emu_paths) as run:
        ref, got = run(ce)
    tol = 1e-4 if dtype == torch.float32 else 2e-2
    close(got[0], ref[0], rtol=tol, atol_frac=tol)
    close(got[1], ref[1], rtol=tol * 10, atol_frac=tol)
    close(got[2], ref[2], rtol=tol, atol_frac=tol)


# ------------------------------------------------------------------------------------------------------------ BN / GAP
@pytest.mark.parametrize("relu,residual", [(False, False), (True, True)])
def test_batch_norm_nhwc_forward_backward_and_eval(emu_paths, relu, residual):
    m, c = 8 * 16 * 8, 256

    def bn():
        from flpr_b200.ops import fused as fops
        x = rand(m, c, seed=1).to(torch.bfloat16).requires_grad_(True)
        res = rand(m, c, seed=2).to(torch.bfloat16).requires_grad_(True) if residual else None
        gamma = (1 + 0.1 * rand(c, seed=3)).requires_grad_(True)
        beta = (0.1 * rand(c, seed=4)).requires_grad_(True)
        rm, rv = torch.zeros(c), torch.ones(c)
        y = fops.batch_norm_nhwc(x, gamma, beta, rm, rv, training=True, relu=relu, residual=res)
        y.float().mul(rand(m, c, seed=5)).sum().backward()
        ye = fops.batch_norm_nhwc(x.detach(), gamma.detach(), beta.detach(), rm, rv, training=False, relu=relu,
                                  residual=None if res is None else res.detach())
        return y.detach(), x.grad, gamma.grad, beta.grad, rm, rv, ye, (None if res is None else res.grad)
    with both(emu_paths) as run:
        ref, got = run(bn)
    close(got[0], ref[0], rtol=2e-2, atol_frac=1e-2)
    close(got[1], ref[1], rtol=5e-2, atol_frac=2e-2)
    close(got[2], ref[2], rtol=2e-2, atol_frac=1e-2)
    close(got[3], ref[3], rtol=2e-2, atol_frac=1e-2)
    close(got[4], ref[4], rtol=1e-3, atol_frac=1e-3)
    close(got[5], ref[5], rtol=1e-3, atol_frac=1e-3)
    close(got[6], ref[6], rtol=2e-2, atol_frac=1e-2)
    if residual:
        close(got[7], ref[7], rtol=2e-2, atol_frac=1e-2)


def test_global_average_pool(emu_paths):
    def gap():
        from flpr_b200.ops import fused as fops
        x = rand(6, 32, 128, seed=1).to(torch.bfloat16).requires_grad_(True)
        y = fops.global_avg_pool_nhwc(x)
        y.float().mul(rand(6, 128, seed=2)).sum().backward()
        return y.detach(), x.grad
    with both(emu_paths) as run:
        ref, got = run(gap)
    close(got[0], ref[0], rtol=1e-2, atol_frac=1e-2)
    close(got[1], ref[1], rtol=1e-2, atol_frac=1e-2)


# ------------------------------------------------------------------------------------------------------------ attention
@pytest.mark.parametrize("n,heads,d,nw", [(16, 3, 32, 1), (49, 2, 32, 4)])
def test_scalar_window_attention_forward_backward(emu_paths, n, heads, d, nw):
    bw = 2 * nw

    def attn():
        from flpr_b200.ops import fused as fops
        qkv = rand(bw, n, 3, heads, d, seed=1, scale=0.5).requires_grad_(True)          # fp32: the scalar kernel
        bias = rand(nw, heads, n, n, seed=2, scale=0.2).requires_grad_(True)
        out = fops.window_attention(qkv, bias, d ** -0.5)
        out.mul(rand(bw, n, heads * d, seed=3)).sum().backward()
        return out.detach(), qkv.grad, bias.grad
    with both(emu_paths) as run:
        ref, got = run(attn)
    close(got[0], ref[0], rtol=1e-3, atol_frac=1e-4)
    close(got[1], ref[1], rtol=1e-3, atol_frac=1e-3)
    close(got[2], ref[2], rtol=1e-3, atol_frac=1e-3)


# ------------------------------------------------------------------------------------------------------------ losses
@pytest.mark.parametrize("norm_feat,hard,margin", [(False, True, 0.3), (True, False, None), (False, False, 0.3),
                                                   (True, True, None)])
def test_triplet_mining_kernels(emu_paths, norm_feat, hard, margin):
    def trip():
        from flpr_b200.criterions.losses import TripletLoss
        x = rand(32, 256, seed=1).requires_grad_(True)
        y = torch.arange(32) // 4
        loss = TripletLoss(margin=margin, norm_feat=norm_feat, hard_mining=hard)(feature=x, target=y)
        loss.backward()
        return loss.detach(), x.grad
    with both(emu_paths) as run:
        ref, got = run(trip)
    close(got[0], ref[0], rtol=2e-2, atol_frac=2e-2)           # (the fused path forms the Gram matrix in bf16)
    cos = F.cosine_similarity(got[1].flatten(), ref[1].flatten(), dim=0).item()
    assert cos > 0.99, cos


def test_kd_kl_and_bce_distill(emu_paths):
    def kd():
        from flpr_b200.ops import fused as fops
        s = rand(16, 500, seed=1, scale=2.0).requires_grad_(True)
        t = rand(16, 500, seed=2, scale=2.0)
        loss = fops.kd_kl(s, t, 4.0)
        loss.backward()
        z = rand(16, 40, seed=3).requires_grad_(True)
        target = torch.randint(0, 40, (16,), generator=torch.Generator().manual_seed(4))
        prev = rand(16, 30, seed=5)
        l2 = fops.bce_distill(z, target, prev)
        l2.backward()
        return loss.detach(), s.grad, l2.detach(), z.grad
    with both(emu_paths) as run:
        ref, got = run(kd)
    for a, b in zip(got, ref):
        close(a, b, rtol=1e-3, atol_frac=1e-4)


# ------------------------------------------------------------------------------------------------------------ the rest
def test_ranking_kernel_cmc_and_map(emu_paths):
    """``rank_eval_kernel`` (count-based AP / first-match rank per query, ``tools/evaluate.py:103-142``) vs the reference
    formulation, incl. queries without any match and tied similarities."""
    def rank():
        from flpr_b200.ops.rank import rank_metrics
        g = torch.Generator().manual_seed(5)
        sim = torch.randn(37, 211, generator=g)
        sim[:, 0:210:7] = sim[:, 1:210:7]                                    # ties
        ql = torch.randint(0, 12, (37,), generator=g)
        gl = torch.randint(0, 10, (211,), generator=g)                       # ids 10, 11 never appear in the gallery
        cmc, mAP = rank_metrics(sim, ql, gl)
        return torch.as_tensor(cmc), torch.tensor(mAP)
    with both(emu_paths) as run:
        ref, got = run(rank)
    close(got[0], ref[0], rtol=1e-6, atol_frac=1e-7)
    close(got[1], ref[1], rtol=1e-5, atol_frac=1e-6)


def test_herding_kernel_picks_the_same_exemplars(emu_paths):
    def herd():
        from flpr_b200.methods.fedstil import group_matrix, herding_select_batched
        g = torch.Generator().manual_seed(9)
        feats = torch.randn(60, 48, generator=g)
        groups = [torch.arange(0, 17), torch.arange(17, 20), torch.arange(20, 60)]
        idx, counts = group_matrix(groups)
        return (herding_select_batched(feats, idx, counts, 6),)
    with both(emu_paths) as run:
        ref, got = run(herd)
    counts = [17, 3, 40]
    for gi, c in enumerate(counts):                        # (entries past an identity's size are padding on both paths)
        k = min(6, c)
        assert torch.equal(got[0][gi, :k], ref[0][gi, :k]), (gi, got[0][gi], ref[0][gi])


def test_stem_space_to_depth_and_maxpool(emu_paths):
    def stem():
        from flpr_b200.ops import gemm as gops
        x = rand(3, 16, 8, 3, seed=1).to(torch.bfloat16)
        y = rand(2, 9, 7, 64, seed=2).to(torch.bfloat16)
        return gops.s2d_pad(x), gops.maxpool3x3s2(y)
    with both(emu_paths) as run:
        ref, got = run(stem)
    assert torch.equal(got[0], ref[0])
    assert torch.equal(got[1], ref[1])


@pytest.mark.parametrize("level", ["none", "default"])
def test_fused_augmentation_kernel(emu_paths, level):
    """``augment_u8_kernel`` (normalise + flip + random erasing + cast in one pass) consumes the same block of uniforms as
    the tensor-op reference: same generator state -> same batch (``datasets/image_augmentation.py:6-71``)."""
    def aug():
        from flpr_b200.data.augmentation import DeviceAugment
        u8 = torch.randint(0, 256, (12, 32, 16, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(3))
        out = DeviceAugment(level, dtype=torch.float32)(u8, torch.Generator().manual_seed(77))
        return (out.contiguous(),)
    with both(emu_paths) as run:
        ref, got = run(aug)
    assert got[0].shape == ref[0].shape
    frac = float(((got[0] - ref[0]).abs() > 1e-4).float().mean())
    assert frac < 0.02, frac          # an erase rectangle may differ by one border pixel row (round-half cases), not more


# ------------------------------------------------------------------------------------------------------------ integration
def _run_tiny(tmp, method, emu_paths):
    from helpers import tiny_common, tiny_experiment, tiny_factory
    from flpr_b200.ops import native
    from flpr_b200.runtime.experiment import ExperimentStage
    native.use_emulated_libraries(emu_paths)
    try:
        common = tiny_common(str(tmp))
        common["defaults"]["exp_opts"].update(comm_rounds=2, val_interval=2)
        cfg = tiny_experiment(common, method)
        cfg["engine_opts"].update(val_at_round0=False)
        before = native.launches()
        with ExperimentStage(common, [cfg], source_factory=tiny_factory()) as stage:
            log = stage.run_experiment(cfg)
        launches = native.launches() - before
    finally:
        native.use_emulated_libraries(None)
    out = {}
    for client, per_round in log.records["data"].items():
        for rnd, tasks in per_round.items():
            for task, v in tasks.items():
                for key in ("tr_loss", "val_map", "val_rank_1"):
                    if key in v:
                        out[(client, int(rnd), task, key)] = float(v[key])
    return out, launches


@pytest.mark.parametrize("method", ["fedavg", "ewc"] + (["fedstil"] if __import__("os").environ.get("FLPR_EMU_FULL") else []))
def test_whole_experiment_on_the_emulated_kernels(tmp_path, emu_paths, method):
    """A tiny experiment run twice through the engine - PyTorch reference paths, then with the emulated libraries installed,
    so that the fused optimizer, importance accumulation, label-smoothing CE (+ its device accumulators), the augmentation
    kernel, the ranking kernel of the validation pass (and, with ``FLPR_EMU_FULL=1``, FedSTIL's herding / compose) run
    from the engine's own call sites. Logged training losses and CMC / mAP must agree."""
    ref, n0 = _run_tiny(tmp_path / "ref", method, None)
    got, n1 = _run_tiny(tmp_path / "emu", method, emu_paths)
    assert n0 == 0 and n1 >= 20, (n0, n1)
    assert ref.keys() == got.keys() and len(ref) >= 6
    for key in ref:
        assert abs(ref[key] - got[key]) <= 5e-3 * abs(ref[key]) + 1e-6, (key, ref[key], got[key])
