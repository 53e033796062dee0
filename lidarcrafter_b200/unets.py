"""String registry with the reference's keys (lidargen/models/unets/__init__.py:22-37):
``__all__[cfg.model.architecture](in_channels=..., resolution=..., **cfg.model.params)``."""
from .efficient_unet import EfficientUNet
from .layout_encoder import LayoutTransformerEncoder
from .layout_unet_v1 import LayoutUnetV1

__all__ = {
    "efficient_unet": EfficientUNet,
    "layout_unet_v1": LayoutUnetV1,
    "layout_encoder": LayoutTransformerEncoder,
}
