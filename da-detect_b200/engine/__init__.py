from .trainer import FlatSGDTrainer, WarmupMultiStepLR, WarmupCosineLR
