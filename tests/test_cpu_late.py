"""CPU runs of ``late_checks`` (reference paths of the ops: validates the autograd glue, layouts and index maths)."""
import os

import pytest

import late_checks as L


@pytest.mark.parametrize("k,n,h,w,cin,cout", [(3, 4, 16, 8, 128, 64), (1, 4, 16, 8, 128, 256), (3, 2, 32, 16, 64, 128),
                                             (1, 2, 8, 4, 64, 64)])
def test_conv_stride2_matches_conv2d(k, n, h, w, cin, cout):
    L.check_conv_stride2("cpu", k, n, h, w, cin, cout)
    L.check_conv_stride2("cpu", k, n, h, w, cin, cout, want_stats=True)


def test_fast_head_last_stride2():
    L.check_fast_head_last_stride2("cpu")


@pytest.mark.parametrize("channels_last", [False, True])
def test_compose_function_reference_path(channels_last):
    L.check_compose_function("cpu", channels_last, use_ref=True)


@pytest.mark.parametrize("channels_last", [False, True])
def test_fedweit_layer_fused_theta(channels_last):
    L.check_fedweit_layer_fused("cpu", channels_last, use_ref=True)


def test_atten_composer_storage_alignment():
    L.check_atten_composer_storage("cpu", use_ref=True)


def test_swin_token_ops_reference_index_maths():
    L.check_swin_token_ops("cpu")


def test_swin_block_fused_forward():
    L.check_swin_block_fused("cpu")


def test_layer_norm_rows_reference_path():
    L.check_layer_norm_rows("cpu")


def test_swin_train_block_norm_classes():
    L.check_swin_train_block_norms("cpu")


def test_gelu_act_reference_path():
    L.check_gelu_act("cpu")


def test_window_merge_residual_reference_path():
    L.check_window_merge_residual("cpu")


def test_swin_train_block_fused_path_matches_plain_block():
    L.check_swin_train_block_fused("cpu")


def test_apply_global_reference_semantics():
    L.check_apply_global("cpu")


def test_convergence_harness_is_deterministic_on_cpu(tmp_path):
    """Same engine twice (fp32 CPU): the curves the GPU convergence test compares must be reproducible."""
    ref, got = L.check_bf16_engine_tracks_fp32(str(tmp_path), "fedstil", "cpu", per_round=1e-5, mean_tol=1e-5, rounds=3)
    assert all(len(v) == 3 for v in ref.values())


def test_selfcheck_children_protocol(tmp_path, monkeypatch):
    """The isolated self-check plumbing (one child process per kernel family, verdict parsed from its stdout, cached in
    a file) - exercised on the CPU reference paths, where every check must pass."""
    import tempfile
    import torch
    from flpr_b200.ops import layer as lops
    monkeypatch.setattr(tempfile, "tempdir", str(tmp_path))
    verdict = lops._checks_isolated(torch.device("cpu"))
    assert verdict == {"wcompose": True, "swin_tokens": True, "apply": True, "ln_train": True}, verdict
    cached = [f for f in os.listdir(tmp_path) if f.startswith("flpr_layer_selfcheck_")]
    assert len(cached) == 1
    with open(os.path.join(tmp_path, cached[0]), "w") as f:          # the cache is what the next process reads
        f.write('{"wcompose": true, "swin_tokens": false, "apply": true}')
    assert lops._checks_isolated(torch.device("cpu"))["swin_tokens"] is False
    assert lops.run_checks_inprocess("cpu", ["apply"]) == {"apply": True}
