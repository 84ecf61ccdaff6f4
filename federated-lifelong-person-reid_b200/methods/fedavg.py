"""``fedavg`` – FedAvg (reference ``methods/fedavg.py``): upload = all trainable parameters + ``train_cnt``;
server = ``train_cnt``-weighted mean over every registered client's last upload; dispatch = trainable parameters
(full ``state_dict`` on first contact). Aggregation + broadcast run as one fused NVLink kernel (see fedbase)."""
from __future__ import annotations

from ..runtime.modules import OperatorModule
from .fedbase import FedClient, FedServer


class Operator(OperatorModule):
    pass


class Client(FedClient):
    default_ckpt_name = "fedavg_model"


class Server(FedServer):
    pass
