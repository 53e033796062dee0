"""LiDARUtility and ray-angle grids (lidargen/utils/lidar.py:22-132) on the B200 library."""
from __future__ import annotations

import math

import torch
from torch import nn

from . import _lib


def get_linear_ray_angles(H: int = 64, W: int = 2048, fov_up=10, fov_down=-30, device="cpu") -> torch.Tensor:
    """lidar.py:22-32 -- [1,2,H,W] (elevation, azimuth) in radians."""
    elevation = (1 - torch.arange(H, device=device) / H) * (fov_up - fov_down) + fov_down
    azimuth = (1 - torch.arange(W, device=device) / W) * 360.0 - 180.0
    elevation, azimuth = torch.meshgrid([elevation, azimuth], indexing="ij")
    return torch.stack([elevation, azimuth])[None].deg2rad()


class LiDARUtility(nn.Module):
    """lidar.py:34-132.  ``to_xyz_from_normalized`` fuses denormalize -> revert_depth -> to_xyz
    (the post-processing chain of tools/generate/generate.py:51-57) into one kernel."""

    def __init__(self, resolution, depth_format: str, min_depth: float, max_depth: float,
                 ray_angles: torch.Tensor = None):
        super().__init__()
        assert depth_format in ("log_depth", "inverse_depth", "depth")
        if ray_angles is None:
            raise NotImplementedError
        assert ray_angles.ndim == 4 and ray_angles.shape[1] == 2
        self.resolution = tuple(resolution)
        self.depth_format = depth_format
        self.min_depth, self.max_depth = min_depth, max_depth
        ray_angles = torch.nn.functional.interpolate(ray_angles, size=self.resolution, mode="nearest-exact")
        self.register_buffer("ray_angles", ray_angles.float())

    @staticmethod
    def denormalize(x: torch.Tensor) -> torch.Tensor:
        return (x + 1) / 2

    @staticmethod
    def normalize(x: torch.Tensor) -> torch.Tensor:
        return x * 2 - 1

    def get_mask(self, metric):
        return ((metric > self.min_depth) & (metric < self.max_depth)).float()

    @torch.no_grad()
    def to_xyz_from_normalized(self, x_norm: torch.Tensor):
        """x_norm [B,1,H,W] in [-1,1] (sampler output, channel 0) -> (metric depth [B,1,H,W], xyz [B,3,H,W])."""
        if self.depth_format != "log_depth":
            raise NotImplementedError("fused path implements the nuScenes 'log_depth' format")
        assert x_norm.is_cuda and x_norm.dim() == 4 and x_norm.shape[1] == 1
        B, _, H, W = x_norm.shape
        x = x_norm.contiguous().float()
        depth = torch.empty(B, 1, H, W, device=x.device)
        xyz = torch.empty(B, 3, H, W, device=x.device)
        _lib.get_lib().depth_to_xyz(x.data_ptr(), self.ray_angles.data_ptr(), depth.data_ptr(), xyz.data_ptr(), B, H, W,
                                    float(self.min_depth), float(self.max_depth),
                                    torch.cuda.current_stream(x.device).cuda_stream)
        return depth, xyz

    # elementwise pre-processing helpers (used once per sample, torch is fine here: not on the step path)
    @torch.no_grad()
    def convert_depth(self, metric, mask=None, depth_format=None):
        depth_format = depth_format or self.depth_format
        mask = self.get_mask(metric) if mask is None else mask
        if depth_format == "log_depth":
            normalized = torch.log2(metric + 1) / math.log2(self.max_depth + 1)
        elif depth_format == "inverse_depth":
            normalized = self.min_depth / metric.add(1e-8)
        else:
            normalized = metric.div(self.max_depth)
        return normalized.clamp(0, 1) * mask

    @torch.no_grad()
    def revert_depth(self, normalized, image_format=None):
        image_format = image_format or self.depth_format
        if image_format == "log_depth":
            metric = torch.exp2(normalized * math.log2(self.max_depth + 1)) - 1
        elif image_format == "inverse_depth":
            metric = self.min_depth / normalized.add(1e-8)
        else:
            metric = normalized.mul(self.max_depth)
        return metric * self.get_mask(metric)

    @torch.no_grad()
    def to_xyz(self, metric):
        assert metric.dim() == 4
        mask = self.get_mask(metric)
        phi, theta = self.ray_angles[:, [0]], self.ray_angles[:, [1]]
        xyz = torch.cat((metric * phi.cos() * theta.cos(), metric * phi.cos() * theta.sin(), metric * phi.sin()), dim=1)
        return xyz * mask
