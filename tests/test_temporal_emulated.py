"""Temporal (autoregressive 4D) glue on the CPU through the C-ABI emulator, against goldens of the reference's own
pipe_related / custom_dataset / convert_boxes_to_2d functions (tests/golden/make_golden_temporal.py)."""
import numpy as np
import pytest

import temporal_checks as TC
from abi_emulator import EmulatedLib
from lidarcrafter_b200 import _lib
from lidarcrafter_b200 import layout_ops as LO


@pytest.fixture()
def emu():
    lib = EmulatedLib()
    _lib.set_test_lib(lib)
    yield lib
    _lib.set_test_lib(None)


def test_layout_item_matches_reference_dataset_item(emu):
    TC.check_layout_item()
    assert "boxes_to_mask" in emu.calls


def test_clip_glue_matches_reference(emu):
    TC.check_clip_glue("cpu")
    assert "range_project_f64" in emu.calls and "points_in_boxes" in emu.calls


def test_batched_boxes_equal_single_calls(emu):
    """[F,N,8] batches (all frames of a clip in one launch) give the per-frame results"""
    G = np.load(TC.os.path.join(TC.os.path.dirname(TC.os.path.abspath(__file__)), "golden", "boxes2d.npz"))
    for dt, cases in ((np.float64, (0, 1, 2)), (np.float32, (3, 4, 5))):
        stack = np.stack([G[f"boxes_{c}"] for c in cases]).astype(dt)
        b2, mask, w = LO.convert_boxes_to_2d(stack, H=32, W=1024, fov_up=10.0, fov_down=-30.0)
        for i, c in enumerate(cases):
            assert np.array_equal(b2[i], G[f"b2_{c}"]) and np.array_equal(mask[i], G[f"mask_{c}"])
            assert np.allclose(w[i], G[f"w_{c}"], rtol=1e-6)
