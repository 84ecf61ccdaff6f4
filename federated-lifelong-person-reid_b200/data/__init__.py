from .augmentation import augmentations, DeviceAugment  # noqa: F401
from .datasets import ReIDImageDataset, ArrayReIDDataset  # noqa: F401
from .pipeline import ReIDTaskPipeline, DeviceBatchLoader  # noqa: F401
