"""Forgetting analysis (``analyse/forgetting.py``): for every task, the best value seen so far minus every later
value, averaged over (task, later round) pairs per client and then over clients."""
from __future__ import annotations

from typing import Dict, List, Tuple

from .accuracy import _plt


def forgetting_on_round(logs: Dict, rounds: int, metric: str, metric_desc: str = "", verbose: bool = True
                        ) -> Tuple[Dict[str, float], float]:
    per_client: Dict[str, float] = {}
    for client_name, communication in logs.items():
        best: Dict[str, Tuple[float, int]] = {}
        for rnd in sorted(communication, key=int):
            r = int(rnd)
            if r > rounds:
                break
            for task, vals in communication[rnd].items():
                if metric in vals and (task not in best or vals[metric] > best[task][0]):
                    best[task] = (vals[metric], r)
        drops: List[float] = []
        for task, (value, r) in best.items():
            for later in range(r + 1, rounds + 1):
                entry = communication.get(str(later), {}).get(task, {})
                if metric in entry:
                    drops.append(value - entry[metric])
        if drops:
            per_client[client_name] = sum(drops) / len(drops)
            if verbose:
                print(f"[{client_name}] {metric} has forgetting {per_client[client_name]:.2%}")
    total = sum(per_client.values()) / max(len(per_client), 1)
    if verbose:
        print(f"Total clients {metric_desc or metric} has forgetting {total:.2%}.")
    return per_client, total


def forgetting_curves(logs: Dict, metric: str) -> List[Tuple[int, float]]:
    rounds = sorted({int(r) for comm in logs.values() for r, tasks in comm.items()
                     if any(metric in v for v in tasks.values())})
    return [(r, forgetting_on_round(logs, r, metric, verbose=False)[1]) for r in rounds]


def plot_forgetting_for_many_jobs(jobs: Dict[str, Dict], save_path_prefix: str, metric: str, metric_desc: str = ""
                                  ) -> None:
    plt = _plt()
    clients = sorted({c for logs in jobs.values() for c in logs})
    for client in clients:
        fig, ax = plt.subplots(figsize=(6, 4))
        for job, logs in jobs.items():
            if client in logs:
                pts = forgetting_curves({client: logs[client]}, metric)
                ax.plot([p[0] for p in pts], [p[1] for p in pts], marker="o", label=job)
        ax.set_title(client); ax.set_xlabel("communication round"); ax.set_ylabel(f"forgetting of {metric_desc or metric}")
        ax.legend(); fig.tight_layout(); fig.savefig(f"{save_path_prefix}_{client}.png"); plt.close(fig)


def plot_merged_forgetting_for_many_jobs(jobs: Dict[str, Dict], save_path: str, metric: str, metric_desc: str = ""
                                         ) -> None:
    plt = _plt()
    fig, ax = plt.subplots(figsize=(6, 4))
    for job, logs in jobs.items():
        pts = forgetting_curves(logs, metric)
        ax.plot([p[0] for p in pts], [p[1] for p in pts], marker="o", label=job)
    ax.set_xlabel("communication round"); ax.set_ylabel(f"forgetting of {metric_desc or metric}"); ax.legend()
    fig.tight_layout(); fig.savefig(save_path); plt.close(fig)
