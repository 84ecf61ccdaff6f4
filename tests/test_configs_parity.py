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


def test_batch_launcher_runs_the_reference_experiment_list():
    """``startup.sh`` (reference: ``startup.sh``): valid shell, launches ``main.py`` in the background with the same ten
    basis experiments in the same order, each of which exists here."""
    import re
    import subprocess
    mine = os.path.join(ROOT, "startup.sh")
    assert subprocess.run(["bash", "-n", mine]).returncode == 0

    def experiments(path):
        with open(path) as f:
            return [os.path.normpath(p) for p in re.findall(r"[\w./]*configs/[\w./]+\.yaml", f.read())]

    ours = experiments(mine)
    assert len(ours) == 10 and all(os.path.exists(os.path.join(ROOT, p)) for p in ours)
    ref_sh = "/root/reference/startup.sh"               # (not part of the pip-installed copy in baseline/_ref)
    if os.path.exists(ref_sh):
        assert ours == experiments(ref_sh)
    with open(mine) as f:
        text = f.read()
    assert "main.py --experiments" in text and "nohup" in text and text.rstrip().endswith("&")
