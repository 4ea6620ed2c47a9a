"""CPU: the inference execution plan of the encoders (BN folding, SE gate folded into the projection GEMM, padded
staging buffers) computes the same function as the exact EfficientNet-B0 module (fp32, tolerance 1e-4 of max|ref|)."""
import pytest
import torch

from ccvpe_b200.efficientnet import EfficientNetB0
from ccvpe_b200.fast_encoder import FastEncoder
from ccvpe_b200.synthetic import fill_deterministic
from helpers import rel_err


@pytest.mark.parametrize("circular,shape", [(True, (2, 3, 64, 128)), (False, (1, 3, 96, 96)), (False, (1, 3, 77, 115))])
def test_fast_encoder_matches_exact_fp32(circular, shape):
    enc = EfficientNetB0(circular=circular).eval()
    fill_deterministic(enc.state_dict(), seed=5)
    fast = FastEncoder(enc, dtype=torch.float32)
    x = torch.randn(*shape, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        ref_head, ref_blocks = enc.extract_features_multiscale(x)
    for _ in range(2):                                   # second call reuses the cached padded buffers
        head, blocks = fast.extract_features_multiscale(x)
        assert head.shape == ref_head.shape
        assert rel_err(head, ref_head) < 1e-4
        assert len(blocks) == 16
        for a, b in zip(blocks, ref_blocks):
            assert a.shape == b.shape and rel_err(a, b) < 1e-4
    assert rel_err(fast.extract_features(x), ref_head) < 1e-4
