from .trainer import FlatSGDTrainer, WarmupMultiStepLR, WarmupCosineLR
from .prefetch import DevicePrefetcher
