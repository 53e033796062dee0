"""convert_boxes_to_2d / preprocessors against goldens of the unmodified reference
(lidargen/dataset/transforms_3d/common.py:99-181; tests/golden/make_golden_temporal.py writes boxes2d.npz)."""
import os

import numpy as np
import pytest
import torch

import lidarcrafter_b200 as L
from abi_emulator import EmulatedLib
from lidarcrafter_b200 import _lib
from lidarcrafter_b200 import layout_ops as LO

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "boxes2d.npz"))


@pytest.fixture()
def emu():
    lib = EmulatedLib()
    _lib.set_test_lib(lib)
    yield lib
    _lib.set_test_lib(None)


def test_convert_boxes_to_2d_without_the_library_fails_loudly():
    """the product path has no host fallback: on a machine without a B200 the call raises"""
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(Exception):
        LO.convert_boxes_to_2d(G["boxes_0"], H=32, W=1024, fov_up=10.0, fov_down=-30.0)


def test_convert_boxes_to_2d_matches_reference(emu):
    """through the C-ABI emulator (= the C oracle): float64 (cases 0-2) and float32 (3-5) boxes, seam-straddling boxes"""
    for case in range(6):
        b2, mask, w = LO.convert_boxes_to_2d(G[f"boxes_{case}"], H=32, W=1024, fov_up=10.0, fov_down=-30.0)
        assert np.array_equal(b2, G[f"b2_{case}"])
        assert np.array_equal(mask, G[f"mask_{case}"])
        assert np.allclose(w, G[f"w_{case}"], rtol=1e-6)


def test_preprocess_condition_mask_shapes_and_values():
    lu = L.LiDARUtility((32, 1024), "log_depth", 1.45, 80.0, L.get_linear_ray_angles(32, 1024, 10, -30))
    cm = torch.from_numpy(G["mask_0"])[None]
    cc = LO.preprocess_condition_mask(cm, lu)
    assert cc.shape == (1, 10, 32, 1024)
    assert torch.equal(cc[:, :9].sum(1), torch.ones(1, 32, 1024))
    d = cm[:, 1]
    ref = (torch.log2(d + 1) / np.log2(81.0)).clamp(0, 1) * ((d > 1.45) & (d < 80.0)).float()
    assert torch.allclose(cc[:, 9], ref)
