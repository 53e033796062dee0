"""lidarcrafter_b200 -- B200-native (sm_100a) implementation of the LiDARCrafter denoiser hot path.

Host side mirrors the reference interfaces (string registries, ctor kwargs, state_dict keys,
``GaussianDiffusion.sample``); the math runs in libb200lidar.so (csrc/, C-ABI in include/b200lidar.h).
"""
from ._lib import B200LidarError, LIB_PATH, get_lib, require_b200  # noqa: F401
from .efficient_unet import EfficientUNet  # noqa: F401
from .layout_unet_v1 import LayoutUnetV1  # noqa: F401
from .layout_encoder import LayoutTransformerEncoder  # noqa: F401
from .diffusion import (ContinuousTimeGaussianDiffusion, CondContinuousTimeGaussianDiffusion,  # noqa: F401
                        GaussianDiffusion)
from .lidar import LiDARUtility, get_linear_ray_angles  # noqa: F401
from . import unets  # noqa: F401

__all__ = ["EfficientUNet", "LayoutUnetV1", "LayoutTransformerEncoder", "ContinuousTimeGaussianDiffusion",
           "CondContinuousTimeGaussianDiffusion", "GaussianDiffusion", "LiDARUtility",
           "get_linear_ray_angles", "unets", "get_lib", "require_b200", "B200LidarError"]
