"""Micro-benchmark of the fused bias+SiLU(+channel-sum) encoder kernel (development tool)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ccvpe_b200 import cabi

dev = torch.device("cuda")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
for (B, H, W, C, padded, sums) in [(64, 256, 256, 96, True, False), (64, 128, 128, 96, False, True), (64, 128, 128, 144, False, True),
                                   (64, 160, 320, 96, True, False), (64, 32, 32, 480, False, True), (64, 16, 16, 1152, False, True)]:
    x = torch.randn(B, H, W, C, device=dev, dtype=torch.bfloat16)
    bias = torch.randn(C, device=dev, dtype=torch.bfloat16)
    if padded:
        buf = torch.zeros(B, H + 1, W + 1, C, device=dev, dtype=torch.bfloat16)
        out = buf[:, :H, :W, :]
    else:
        out = torch.empty_like(x)
    s = torch.zeros(B, C, device=dev, dtype=torch.int64) if sums else None
    for _ in range(3):
        cabi.bias_silu_nhwc(x, bias, out, s)
    torch.cuda.synchronize()
    ms = []
    for _ in range(8):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); cabi.bias_silu_nhwc(x, bias, out, s); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    best = min(ms)
    print("bias_silu B=%d %dx%dx%d padded=%s sums=%s: %.4f ms  %.1f GB/s" % (B, H, W, C, padded, sums, best, 2 * x.numel() * 2 / best / 1e6))
