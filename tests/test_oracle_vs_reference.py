"""CPU, authoring container only: the oracle against the LIVE reference imported from /root/reference.
Skipped wherever the reference checkout is absent (e.g. on the GPU box)."""
import os

import pytest
import torch

from helpers import build_model, oracle_forward
from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="/root/reference not present")


@pytest.mark.parametrize("variant,shape,noise,circular", [
    ("vigor", (320, 640), None, True),
    ("vigor_prior", (320, 192), 36.0, False),
    ("kitti", (256, 1024), None, None),
    ("oxford", (154, 231), None, None),
])
def test_bit_exact_against_live_reference(variant, shape, noise, circular):
    ref_models = ref_shim.load_reference_models()
    mine = build_model(variant, noise, circular, wseed=21)
    ctor = {"vigor": lambda: ref_models.CVM_VIGOR("cpu", circular),
            "vigor_prior": lambda: ref_models.CVM_VIGOR_ori_prior("cpu", noise, circular),
            "kitti": lambda: ref_models.CVM_KITTI("cpu"),
            "oxford": lambda: ref_models.CVM_OxfordRobotCar("cpu")}[variant]
    ref = ctor().eval()
    ref.load_state_dict(mine.state_dict(), strict=True)           # same keys, same shapes
    g = torch.Generator().manual_seed(99)
    grd = torch.randn((1, 3) + shape, generator=g)
    sat = torch.randn((1, 3, 512, 512), generator=g)
    with torch.no_grad():
        out_ref = ref(grd, sat)
    out_or = oracle_forward(mine, variant, noise, grd, sat)       # our encoder + oracle decoder
    for a, b in zip(out_ref, out_or):
        assert a.shape == b.shape
        assert torch.equal(a, b)                                   # same ops, same order, same thread count


def test_oracle_losses_bit_equal_reference_losses():
    """oracle infonce / cross-entropy / orientation losses and the loss combination == the reference's losses.py:4-29 and
    train_VIGOR.py:120-146 (the training-step oracle of tests/test_gpu_train.py is pinned here)."""
    import importlib.util

    from ccvpe_b200.synthetic import synthetic_ground_truth
    from oracle import ccvpe_oracle as orc
    spec = importlib.util.spec_from_file_location("_ccvpe_reference_losses", os.path.join(ref_shim.REFERENCE_ROOT, "losses.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    g = torch.Generator().manual_seed(5)
    s = torch.rand(3, 20 * 64, generator=g) * 2 - 1
    lab = torch.rand(3, 20 * 64, generator=g) * (torch.rand(3, 20 * 64, generator=g) > 0.9)
    assert torch.equal(ref.infoNCELoss(s, lab), orc.infonce_loss(s, lab))
    logits = torch.randn(2, 4096, generator=g)
    soft = torch.softmax(torch.randn(2, 4096, generator=g), dim=1)
    assert torch.equal(ref.cross_entropy_loss(logits, soft), orc.cross_entropy_loss(logits, soft))
    gt, gwo, gor = synthetic_ground_truth(2, seed=1, size=64)
    ori = torch.nn.functional.normalize(torch.randn(2, 2, 64, 64, generator=g), dim=1)
    assert torch.equal(ref.orientation_loss(ori, gor, gt), orc.orientation_loss(ori, gor, gt))
    # the combination of train_VIGOR.py:120-146, restated with the reference's own functions
    outs = [logits.repeat(1, 1)[:, :64 * 64], None, ori] + [torch.rand(2, 20, 64 // k, 64 // k, generator=g) * 2 - 1
                                                             for k in (64, 32, 16, 8, 4, 2)]
    gt_flat = torch.flatten(gt, start_dim=1)
    gt_flat = gt_flat / torch.sum(gt_flat, dim=1, keepdim=True)
    nce = sum(ref.infoNCELoss(torch.flatten(o, start_dim=1), torch.flatten(torch.nn.MaxPool2d(k, stride=k)(gwo), start_dim=1))
              for o, k in zip(outs[3:], (64, 32, 16, 8, 4, 2)))
    want = ref.cross_entropy_loss(outs[0], gt_flat) + 1e4 * nce / 6 + 1e1 * ref.orientation_loss(ori, gor, gt)
    assert torch.allclose(orc.training_loss(outs, gt, gwo, gor), want, rtol=1e-6)
