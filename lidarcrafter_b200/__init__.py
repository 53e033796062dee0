"""lidarcrafter_b200 -- B200-native (sm_100a) implementation of the LiDARCrafter denoiser hot path.

Host side mirrors the reference interfaces (string registries, ctor kwargs, state_dict keys,
``GaussianDiffusion.sample``); the math runs in libb200lidar.so (csrc/, C-ABI in include/b200lidar.h).
"""
from ._lib import B200LidarError, LIB_PATH, get_lib, require_b200  # noqa: F401
from .efficient_unet import EfficientUNet  # noqa: F401
from .diffusion import ContinuousTimeGaussianDiffusion, GaussianDiffusion  # noqa: F401
from .lidar import LiDARUtility, get_linear_ray_angles  # noqa: F401
from . import unets  # noqa: F401

__all__ = ["EfficientUNet", "ContinuousTimeGaussianDiffusion", "GaussianDiffusion", "LiDARUtility",
           "get_linear_ray_angles", "unets", "get_lib", "require_b200", "B200LidarError"]
