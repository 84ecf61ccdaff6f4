"""Shared tiny-experiment fixtures for the CPU end-to-end tests."""
import copy
import os

from flpr_b200.data.synthetic import synthetic_source_factory
from flpr_b200.runtime.config import merge_experiment


def tiny_common(tmp: str, device="cpu"):
    return {"datasets_dir": os.path.join(tmp, "data"), "checkpoints_dir": os.path.join(tmp, "ckpts"),
            "logs_dir": os.path.join(tmp, "logs"), "parallel": 1, "device": [device],
            "defaults": {
                "random_seed": 123,
                "exp_opts": {"comm_rounds": 2, "val_interval": 1, "online_clients": 2},
                "model_opts": {"name": "resnet18", "num_classes": 8000, "last_stride": 1, "neck": "bnneck",
                               "fine_tuning": ["base.layer4", "classifier"]},
                "criterion_opts": {"name": "cross_entropy", "num_classes": 8000, "epsilon": 0.1},
                "optimizer_opts": {"name": "adam", "lr": 1e-3, "weight_decay": 1e-5},
                "scheduler_opts": {"name": "step_lr", "step_size": 5},
                "task_opts": {"sustain_rounds": 1, "train_epochs": 1,
                              "augment_opts": {"level": "default", "img_size": [32, 16],
                                               "norm_mean": [0.485, 0.456, 0.406], "norm_std": [0.229, 0.224, 0.225]},
                              "loader_opts": {"batch_size": 4, "num_workers": 0, "pin_memory": False,
                                              "persistent_workers": False, "multiprocessing_context": None}}}}


METHOD_OPTS = {
    "fedstil": dict(atten_default=0.9, lambda_l1=1e-4, lambda_k=16),
    "fedstil-atten": dict(atten_default=0.5, lambda_l1=1e-5, lambda_k=16),
    "fedweit": dict(lambda_l1=1e-3, lambda_l2=100.0, lambda_mask=0.0, kb_cnt=2),
    "fedprox": dict(lambda_l2=1e-2),
    "fedcurv": dict(lambda_penalty=10.0),
    "ewc": dict(lambda_penalty=50.0),
    "mas": dict(lambda_penalty=0.01),
    "icarl": dict(k=16, n_classes=10),
}


def tiny_experiment(common, method: str, n_clients: int = 2, n_tasks: int = 2, **over):
    exp = {"exp_name": f"t-{method}", "exp_method": method, "server": {"server_name": "server"},
           "clients": [{"client_name": f"client-{i}", "tasks": [f"task-{i}-{t}" for t in range(n_tasks)]}
                       for i in range(n_clients)]}
    if method in METHOD_OPTS:
        exp["model_opts"] = dict(common["defaults"]["model_opts"], **METHOD_OPTS[method])
    if method.startswith("fedstil"):
        exp["server"].update(distance_calculate_step=1, distance_calculate_decay=0.8)
        for c in exp["clients"]:
            c["model_ckpt_name"] = "fedstil_model"
    exp.update(over)
    return merge_experiment(common, exp)


def tiny_factory():
    return synthetic_source_factory(num_ids=3, train_per_id=3, query_per_id=1, gallery_per_id=2, size=(32, 16))
