"""``mas`` – Memory Aware Synapses variant of the reference (``methods/mas.py``): importance is
``sum |grad of the training loss|`` over *all* remembered loaders including the current task, and the remembered
loader is the task's **query** loader (``mas.py:61-75,416``)."""
from __future__ import annotations

from . import ewc
from .penalty import PenaltyModel


class Model(PenaltyModel):
    importance_mode = "mas"
    skip_current_task = False


class Operator(ewc.Operator):
    pass


class Client(ewc.Client):
    default_ckpt_name = "mas_model"
    remember_split = "query"


class Server(ewc.Server):
    pass
