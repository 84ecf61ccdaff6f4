"""Backbone / optimizer / scheduler registries (``models/__init__.py:6-25``)."""
from ..runtime.arena import optimizers, schedulers  # noqa: F401
from .resnet import resnet18, resnet34, resnet50, resnet101, resnet152


def _lazy_swin(name):
    def ctor(**kwargs):
        from . import swin
        return getattr(swin, name)(**kwargs)
    ctor.__name__ = name
    return ctor


nets = {
    "resnet18": resnet18,
    "resnet34": resnet34,
    "resnet50": resnet50,
    "resnet101": resnet101,
    "resnet152": resnet152,
    "swin_transformer_tiny": _lazy_swin("swin_transformer_tiny"),
    "swin_transformer_small": _lazy_swin("swin_transformer_small"),
    "swin_transformer_base": _lazy_swin("swin_transformer_base"),
    "swin_transformer_large": _lazy_swin("swin_transformer_large"),
}
