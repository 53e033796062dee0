"""String registry with the reference's keys (lidargen/models/unets/__init__.py:22-37):
``__all__[cfg.model.architecture](in_channels=..., resolution=..., **cfg.model.params)``."""
from .efficient_unet import EfficientUNet

__all__ = {
    "efficient_unet": EfficientUNet,
}
