"""GPU runs of ``late_checks`` on the native kernels. This file sorts LAST on purpose: the kernels it covers
(``csrc/layer_ops.cu``, the stride-2 backward route) were written after the round's last GPU session, so under
``pytest -x`` a surprise here cannot hide the result of any other test. Each check compares the kernel against a plain
fp32 PyTorch formulation of the same op."""
import pytest
import torch

import late_checks as L

pytestmark = pytest.mark.gpu


def test_layer_kernel_self_checks_pass():
    """The on-device self-checks that gate the fused paths (a failed family falls back to PyTorch with a logged error -
    this test makes that visible instead of silent)."""
    from flpr_b200.ops import layer as lops
    assert lops.enabled("wcompose", "cuda:0"), "compose kernels failed their numerics self-check"
    assert lops.enabled("swin_tokens", "cuda:0"), "Swin token kernels failed their numerics self-check"
    assert lops.enabled("apply", "cuda:0"), "dispatch-apply kernel failed its numerics self-check"
    assert lops.enabled("ln_train", "cuda:0"), "LayerNorm forward / backward kernels failed their numerics self-check"


@pytest.mark.parametrize("k,n,h,w,cin,cout", [(3, 8, 16, 8, 512, 512), (1, 8, 16, 8, 1024, 2048), (3, 4, 32, 16, 128, 256),
                                             (1, 4, 32, 16, 256, 512), (3, 16, 16, 8, 256, 512)])
@pytest.mark.parametrize("want_stats", [False, True])
def test_conv_stride2_native(k, n, h, w, cin, cout, want_stats):
    L.check_conv_stride2("cuda", k, n, h, w, cin, cout, want_stats=want_stats, slot=want_stats)


def test_fast_head_last_stride2_native():
    L.check_fast_head_last_stride2("cuda")


@pytest.mark.parametrize("channels_last", [False, True])
def test_compose_kernels(channels_last):
    L.check_compose_function("cuda", channels_last, use_ref=False)


@pytest.mark.parametrize("channels_last", [False, True])
def test_fedweit_layer_fused_theta_native(channels_last):
    L.check_fedweit_layer_fused("cuda", channels_last, use_ref=False)


def test_atten_composer_native():
    L.check_atten_composer_storage("cuda", use_ref=False)


def test_apply_global_kernel():
    L.check_apply_global("cuda")


def test_swin_token_kernels():
    L.check_swin_token_ops("cuda")


def test_layer_norm_rows_native():
    L.check_layer_norm_rows("cuda")


def test_swin_train_block_norms_native():
    L.check_swin_train_block_norms("cuda")


def test_gelu_act_native():
    L.check_gelu_act("cuda")


def test_window_merge_residual_native():
    L.check_window_merge_residual("cuda")


def test_swin_train_block_fused_native():
    L.check_swin_train_block_fused("cuda")


def test_swin_block_fused_native():
    L.check_swin_block_fused("cuda")


def test_fedweit_and_atten_steps_use_the_compose_kernel(tmp_path):
    """A FedWeIT / fedstil-atten experiment on the GPU launches the fused compose kernels (when their self-check
    passed) and stays finite."""
    from helpers import tiny_common, tiny_experiment
    from flpr_b200.ops import layer as lops
    from flpr_b200.runtime.experiment import ExperimentStage
    import flpr_b200.data.synthetic as syn
    if not lops.enabled("wcompose", "cuda:0"):
        pytest.skip("compose kernels disabled by their self-check (covered by test_layer_kernel_self_checks_pass)")
    calls = {"n": 0}
    orig = lops.wcompose_fwd

    def counted(*a, **k):
        calls["n"] += 1
        return orig(*a, **k)
    lops.wcompose_fwd = counted
    try:
        for method in ("fedweit", "fedstil-atten"):
            before = calls["n"]
            common = tiny_common(str(tmp_path / method), device="cuda:0")
            common["defaults"]["task_opts"]["augment_opts"]["img_size"] = [64, 32]
            common["defaults"]["task_opts"]["loader_opts"]["batch_size"] = 8
            cfg = tiny_experiment(common, method)
            with ExperimentStage(common, [cfg], source_factory=syn.synthetic_source_factory(
                    num_ids=4, train_per_id=4, size=(64, 32))) as stage:
                log = stage.run_experiment(cfg)
            assert calls["n"] > before, f"{method}: the fused compose kernel never ran"
            assert log is not None
    finally:
        lops.wcompose_fwd = orig


@pytest.mark.parametrize("method", ["fedavg", "fedstil"])
def test_bf16_gpu_engine_tracks_fp32_cpu_engine(tmp_path, method):
    """Ten rounds of the same deterministic tiny experiment on the fp32 CPU engine and on the bf16 GPU engine: per-round
    training loss within 10 % (5 % on average)."""
    L.check_bf16_engine_tracks_fp32(str(tmp_path), method, "cuda:0")


def test_fp32_gpu_engine_tracks_fp32_cpu_engine(tmp_path):
    """``engine_opts.compute_dtype: fp32`` on the GPU (``bench.py --dtype fp32``, the precision-matched arm): same
    experiment, library fp32 kernels + the fused fp32 optimizer - within 5 % per round of the CPU engine."""
    L.check_bf16_engine_tracks_fp32(str(tmp_path), "fedstil", "cuda:0", per_round=0.05, mean_tol=0.02, dtype="fp32")


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least two GPUs on the box")
def test_collective_watchdog_fault_injection():
    """A rank that never enters a collective: the survivors' kernels time out (flag watchdog), skip their store phase
    and the host sees a ``NativeError`` - no hang, no partial aggregate (``tests/dist_comm_fault_check.py``)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--standalone", "--local-addr", "127.0.0.1",
                        "--nproc-per-node", "2", os.path.join(root, "tests", "dist_comm_fault_check.py")],
                       capture_output=True, text=True, timeout=600)
    print(r.stdout[-2000:])
    assert r.returncode == 0 and "DIST_COMM_FAULT_CHECK OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
