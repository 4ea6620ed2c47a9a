"""CPU, authoring container only: the oracle against the LIVE reference imported from /root/reference.
Skipped wherever the reference checkout is absent (e.g. on the GPU box)."""
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
