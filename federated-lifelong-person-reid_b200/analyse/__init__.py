"""Offline analysis over the experiment-log JSON (``analyse/`` of the reference): accuracy tables / curves,
forgetting metric, Grad-CAM visualisation. Plotting needs matplotlib (optional dependency); the numeric functions
have no extra dependency and return their results instead of only printing them."""
import json
from typing import Dict

from .accuracy import (accuracy_on_round, accuracy_curves, plot_accuracy_for_one_job,  # noqa: F401
                       plot_accuracy_for_many_jobs, plot_task_accuracy_for_many_jobs,
                       plot_merged_accuracy_for_many_jobs)
from .forgetting import (forgetting_on_round, forgetting_curves, plot_forgetting_for_many_jobs,  # noqa: F401
                         plot_merged_forgetting_for_many_jobs)
from .visualize import grad_cam, visualize_models  # noqa: F401


def load_logs(log_path: str) -> Dict:
    """``data`` sub-tree of an experiment log: ``{client: {round: {task: {metric: value}}}}`` (analyse/__init__.py:8)."""
    with open(log_path, "r") as f:
        return json.load(f)["data"]
