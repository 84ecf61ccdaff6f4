"""Weight initialisers of the ReID strong baseline (``tools/winit.py:8-28``), usable with ``module.apply(...)``."""
from __future__ import annotations

import torch.nn as nn


def weights_init_kaiming(module: nn.Module) -> None:
    if isinstance(module, nn.Linear):
        nn.init.kaiming_normal_(module.weight, a=0, mode="fan_out")
        if module.bias is not None:
            nn.init.zeros_(module.bias)
    elif isinstance(module, (nn.Conv1d, nn.Conv2d, nn.Conv3d)):
        nn.init.kaiming_normal_(module.weight, a=0, mode="fan_in")
        if module.bias is not None:
            nn.init.zeros_(module.bias)
    elif isinstance(module, (nn.BatchNorm1d, nn.BatchNorm2d, nn.BatchNorm3d)) and module.affine:
        nn.init.ones_(module.weight)
        nn.init.zeros_(module.bias)


def weights_init_classifier(module: nn.Module) -> None:
    """Classifier init (std 0.001). The reference tests ``if module.bias:`` which raises on a tensor bias; fixed."""
    if isinstance(module, nn.Linear):
        nn.init.normal_(module.weight, std=0.001)
        if module.bias is not None:
            nn.init.zeros_(module.bias)
