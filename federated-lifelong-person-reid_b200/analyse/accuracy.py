"""Accuracy analysis (``analyse/accuracy.py``): per-round averages and per-job / per-task / merged curves."""
from __future__ import annotations

from typing import Dict, List, Tuple


def _plt():
    try:
        import matplotlib
        matplotlib.use("Agg")
        from matplotlib import pyplot as plt
        return plt
    except Exception as ex:  # pragma: no cover
        raise RuntimeError("plotting needs matplotlib, which is not installed in this environment") from ex


def accuracy_on_round(logs: Dict, rounds: int, metric: str, metric_desc: str = "", verbose: bool = True
                      ) -> Tuple[Dict[str, float], float]:
    """Mean of ``metric`` over the tasks validated at ``rounds`` per client, and the mean over clients."""
    per_client: Dict[str, float] = {}
    for client_name, communication in logs.items():
        vals = [v[metric] for v in communication.get(str(rounds), {}).values() if metric in v]
        if vals:
            per_client[client_name] = sum(vals) / len(vals)
            if verbose:
                print(f"[{client_name}] {metric} is {per_client[client_name]:.2%}")
    total = sum(per_client.values()) / max(len(per_client), 1)
    if verbose:
        print(f"Total clients {metric_desc or metric}:{total:.2%}.")
    return per_client, total


def accuracy_curves(logs: Dict, metric: str) -> Dict[str, Dict[str, List[Tuple[int, float]]]]:
    """``{client: {task: [(round, value), ...]}}`` sorted by round – the data behind every accuracy plot."""
    out: Dict[str, Dict[str, List[Tuple[int, float]]]] = {}
    for client, comm in logs.items():
        for rnd in sorted(comm, key=int):
            for task, vals in comm[rnd].items():
                if metric in vals:
                    out.setdefault(client, {}).setdefault(task, []).append((int(rnd), vals[metric]))
    return out


def merged_curve(logs: Dict, metric: str) -> List[Tuple[int, float]]:
    """Mean over clients of the per-client mean over tasks, for every round that has validation data."""
    rounds = sorted({int(r) for comm in logs.values() for r, tasks in comm.items()
                     if any(metric in v for v in tasks.values())})
    curve = []
    for r in rounds:
        _, total = accuracy_on_round(logs, r, metric, verbose=False)
        curve.append((r, total))
    return curve


def plot_accuracy_for_one_job(logs: Dict, save_path: str, metric: str, metric_desc: str = "") -> None:
    plt = _plt()
    curves = accuracy_curves(logs, metric)
    fig, axes = plt.subplots(1, max(len(curves), 1), figsize=(5 * max(len(curves), 1), 4), squeeze=False)
    for ax, (client, tasks) in zip(axes[0], curves.items()):
        for task, pts in tasks.items():
            ax.plot([p[0] for p in pts], [p[1] for p in pts], marker="o", label=task)
        ax.set_title(client); ax.set_xlabel("communication round"); ax.set_ylabel(metric_desc or metric)
        ax.legend(fontsize=6)
    fig.tight_layout(); fig.savefig(save_path); plt.close(fig)


def plot_accuracy_for_many_jobs(jobs: Dict[str, Dict], save_path_prefix: str, metric: str, metric_desc: str = ""
                                ) -> None:
    plt = _plt()
    clients = sorted({c for logs in jobs.values() for c in logs})
    for client in clients:
        fig, ax = plt.subplots(figsize=(6, 4))
        for job, logs in jobs.items():
            if client in logs:
                pts = merged_curve({client: logs[client]}, metric)
                ax.plot([p[0] for p in pts], [p[1] for p in pts], marker="o", label=job)
        ax.set_title(client); ax.set_xlabel("communication round"); ax.set_ylabel(metric_desc or metric); ax.legend()
        fig.tight_layout(); fig.savefig(f"{save_path_prefix}_{client}.png"); plt.close(fig)


def plot_task_accuracy_for_many_jobs(jobs: Dict[str, Dict], save_path_prefix: str, metric: str,
                                     metric_desc: str = "") -> None:
    plt = _plt()
    per_job = {job: accuracy_curves(logs, metric) for job, logs in jobs.items()}
    keys = sorted({(c, t) for cur in per_job.values() for c, tasks in cur.items() for t in tasks})
    for client, task in keys:
        fig, ax = plt.subplots(figsize=(6, 4))
        for job, cur in per_job.items():
            pts = cur.get(client, {}).get(task)
            if pts:
                ax.plot([p[0] for p in pts], [p[1] for p in pts], marker="o", label=job)
        ax.set_title(f"{client} / {task}"); ax.set_xlabel("communication round"); ax.set_ylabel(metric_desc or metric)
        ax.legend(); fig.tight_layout(); fig.savefig(f"{save_path_prefix}_{client}_{task}.png"); plt.close(fig)


def plot_merged_accuracy_for_many_jobs(jobs: Dict[str, Dict], save_path: str, metric: str, metric_desc: str = ""
                                       ) -> None:
    plt = _plt()
    fig, ax = plt.subplots(figsize=(6, 4))
    for job, logs in jobs.items():
        pts = merged_curve(logs, metric)
        ax.plot([p[0] for p in pts], [p[1] for p in pts], marker="o", label=job)
    ax.set_xlabel("communication round"); ax.set_ylabel(metric_desc or metric); ax.legend()
    fig.tight_layout(); fig.savefig(save_path); plt.close(fig)
