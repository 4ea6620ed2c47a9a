"""Micro-benchmarks of the round-2 encoder kernels through the C ABI (development / profiling tool; inputs > L2).

    python scripts/bench_encoder_ops.py project [B HW mid cout residual]    # MBConv projection (project_tcgen05_kernel)
    python scripts/bench_encoder_ops.py stem [B H W circular u8]            # encoder stem (stem_tcgen05_kernel)
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ccvpe_b200 import cabi  # noqa: E402

dev = torch.device("cuda")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn, reps=8):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return min(ms)


kind = sys.argv[1]
a = [int(v) for v in sys.argv[2:]]
if kind == "project":
    B, HW, mid, cout, res = (a + [64, 12800, 144, 24, 1][len(a):])[:5]
    d = torch.randn(B, HW, mid, device=dev).to(torch.bfloat16)
    wg = (torch.randn(B, cout, mid, device=dev) / mid ** 0.5).to(torch.bfloat16)
    r = torch.randn(B, HW, cout, device=dev).to(torch.bfloat16) if res else None
    bias = torch.randn(cout, device=dev).to(torch.bfloat16)
    out, out2 = torch.empty(B, HW, cout, device=dev, dtype=torch.bfloat16), torch.empty(B, HW, cout, device=dev, dtype=torch.bfloat16)
    t = timeit(lambda: cabi.mbconv_project_nhwc(d, wg, r, out, bias, out2))
    nbytes = (d.numel() + wg.numel() + (r.numel() if res else 0) + 2 * out.numel()) * 2
    print("project B=%d HW=%d mid=%d cout=%d res=%d: %.4f ms  %.0f GB/s (algorithmic, incl. the biased copy)" % (B, HW, mid, cout, res, t, nbytes / t / 1e6))
elif kind == "stem":
    B, H, W, circ, u8 = (a + [64, 512, 512, 0, 0][len(a):])[:5]
    w = (torch.randn(27, 32, device=dev) * 0.2)
    bias = torch.randn(32, device=dev)
    Ho, Wo = (H + 1 - 3) // 2 + 1, (W + 1 - 3) // 2 + 1
    out = torch.zeros(B, Ho + 2, Wo + 2, 32, device=dev, dtype=torch.bfloat16)
    if u8:
        x = torch.randint(0, 256, (B, 3, H, W), device=dev, dtype=torch.uint8)
        t = timeit(lambda: cabi.stem_conv_silu_u8_nhwc(x, w, bias, out, 0, 1, 1, 1, bool(circ), crop_w=W, shift=None))
    else:
        x = torch.randn(B, 3, H, W, device=dev)
        t = timeit(lambda: cabi.stem_conv_silu_nhwc(x, w, bias, out, 0, 1, 1, 1, bool(circ)))
    nbytes = x.numel() * x.element_size() + B * Ho * Wo * 32 * 2
    print("stem B=%d %dx%d circ=%d u8=%d: %.4f ms  %.0f GB/s (algorithmic)" % (B, H, W, circ, u8, t, nbytes / t / 1e6))
