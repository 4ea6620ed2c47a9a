"""CPU: the oracle restatement reproduces the fixtures generated from the UNMODIFIED reference
(oracle/make_golden.py).  Runs everywhere (no /root/reference, no GPU needed)."""
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN_CONFIGS, build_model, check_against_golden, config_inputs, oracle_forward
from oracle import ccvpe_oracle as orc

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", sorted(GOLDEN_CONFIGS))
def test_oracle_matches_reference_fixture(name):
    variant, shape_key, noise, circular, batch, wseed, iseed = GOLDEN_CONFIGS[name]
    golden = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    model = build_model(variant, noise, circular, wseed)
    grd, sat = config_inputs(name)
    out = oracle_forward(model, variant, noise, grd, sat)
    # fp32 CPU, possibly a different thread count than the generating run: tolerance 1e-5 of max-abs per tensor
    check_against_golden(out, golden, tol=1e-5)
    pose = orc.pose_decode(out[1].numpy(), out[2].numpy())
    assert pose["idx"].tolist() == golden["pose.idx"].tolist()          # bit-exact indices
    assert pose["rc"].tolist() == golden["pose.rc"].tolist()
    np.testing.assert_allclose(pose["angle"], golden["pose.angle"], rtol=0, atol=1e-2)
    assert pose["valid"].tolist() == golden["pose.valid"].tolist()


def test_pose_decode_edge_cases():
    """first-occurrence ties, plateau, extremes, invalid (|cos|>1) -- semantics of train_VIGOR.py:297-316."""
    H = W = 8
    heat = np.zeros((5, 1, H, W), np.float32)
    ori = np.zeros((5, 2, H, W), np.float32)
    heat[0, 0, 3, 4] = heat[0, 0, 5, 1] = 1.0          # tie -> first in raster order
    heat[1] = 0.25                                      # plateau -> index 0
    heat[2, 0, 7, 7] = 2.0                              # last element
    heat[3, 0, 0, 0] = 2.0
    heat[4, 0, 2, 2] = 1.0
    ori[0, :, 3, 4] = (0.0, -1.0)
    ori[1, :, 0, 0] = (1.0, 0.0)
    ori[2, :, 7, 7] = (-1.0, 0.0)
    ori[3, :, 0, 0] = (np.float32(0.6), np.float32(0.8))
    ori[4, :, 2, 2] = (1.5, 0.0)                        # invalid
    p = orc.pose_decode(heat, ori)
    assert p["idx"].tolist() == [3 * W + 4, 0, 63, 0, 2 * W + 2]
    assert p["valid"].tolist() == [1, 1, 1, 1, 0]
    np.testing.assert_allclose(p["angle"][:4], [270.0, 0.0, 180.0, np.degrees(np.arccos(np.float64(np.float32(0.6))))])
    assert np.isnan(p["angle"][4])
