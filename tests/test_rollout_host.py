"""Host-side pose arithmetic of the rollout glue against goldens of the unmodified reference
(tools/vis_tools/utils/common.py:116-222)."""
import os

import numpy as np

from lidarcrafter_b200 import rollout

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rollout.npz"))


def test_inter_frame_transforms_match_reference():
    assert np.allclose(rollout.compute_inter_frame_transforms(G["ego"]), G["Ts"], atol=1e-12)


def test_warp_boxes_future_matches_reference():
    assert np.allclose(rollout.warp_boxes_future(G["boxes0"], G["traj_obj"], G["ego"], 0.0), G["fut_boxes"], atol=1e-12)
