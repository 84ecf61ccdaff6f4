"""The shipped experiment configs carry the reference's values (SURVEY §2.1 "Config system": ``common.yaml`` + 46
experiment YAMLs). Compared file by file against ``baseline/_ref/configs`` when the reference is installed."""
import glob
import os

import pytest
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref", "configs")
OURS = os.path.join(ROOT, "configs")
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference is not installed in baseline/_ref")

# keys this engine adds on top of the reference's schema
EXTRA_KEYS = {"engine_opts"}


def _load(path):
    with open(path) as f:
        return yaml.load(f, Loader=yaml.Loader)


def _strip(d):
    return {k: v for k, v in d.items() if k not in EXTRA_KEYS} if isinstance(d, dict) else d


def test_every_reference_experiment_yaml_has_an_identical_counterpart():
    ref_files = sorted(glob.glob(os.path.join(REF, "*", "*.yaml")))
    assert len(ref_files) == 46
    for rf in ref_files:
        rel = os.path.relpath(rf, REF)
        mine = os.path.join(OURS, rel)
        assert os.path.exists(mine), f"missing config {rel}"
        assert _strip(_load(mine)) == _load(rf), rel


def test_common_defaults_match():
    ref, mine = _load(os.path.join(REF, "common.yaml")), _load(os.path.join(OURS, "common.yaml"))
    # the one deliberate difference: the reference ships ``multiprocessing_context: spawn`` with ``num_workers: 0``, which
    # current PyTorch rejects with a ValueError; ours ships ``null``
    assert ref["defaults"]["task_opts"]["loader_opts"].pop("multiprocessing_context") == "spawn"
    assert mine["defaults"]["task_opts"]["loader_opts"].pop("multiprocessing_context") is None
    assert _strip(mine["defaults"]) == ref["defaults"]
    assert {k: v for k, v in mine.items() if k != "defaults"} == {k: v for k, v in ref.items() if k != "defaults"}
