"""Micro-benchmark: fused depthwise conv kernel vs cuDNN depthwise + fused bias/SiLU (development tool)."""
import os, sys
import torch
from torch.nn import functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ccvpe_b200 import cabi

dev = torch.device("cuda")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

def timeit(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(6):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return min(ms)

# all 16 MBConv depthwise layers of EfficientNet-B0: (divisor of the stem resolution, C, K, S)
BLOCKS = [(1, 32, 3, 1), (1, 96, 3, 2), (2, 144, 3, 1), (2, 144, 5, 2), (4, 240, 5, 1), (4, 240, 3, 2), (8, 480, 3, 1),
          (8, 480, 3, 1), (8, 480, 5, 1), (8, 672, 5, 1), (8, 672, 5, 1), (8, 672, 5, 2), (16, 1152, 5, 1), (16, 1152, 5, 1),
          (16, 1152, 5, 1), (16, 1152, 3, 1)]
FAST = len(sys.argv) > 1 and sys.argv[1] == "fast"          # skip the cuDNN comparison
tot = 0.0
cases = [(64, h0 // d, w0 // d, C, K, S) for (h0, w0) in ((160, 320), (256, 256)) for (d, C, K, S) in BLOCKS]
for (B, H, W, C, K, S) in cases:
    lo = (K - S) // 2 if S == 2 else (K - 1) // 2
    hi = (K - S) - lo if S == 2 else (K - 1) // 2
    buf = torch.randn(B, H + lo + hi, W + lo + hi, C, device=dev).to(torch.bfloat16)
    w = (torch.randn(C, 1, K, K, device=dev) * 0.3).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    wt = w.reshape(C, K * K).t().contiguous()
    bias = torch.randn(C, device=dev).to(torch.bfloat16)
    Ho, Wo = (H + lo + hi - K) // S + 1, (W + lo + hi - K) // S + 1
    y = torch.empty(B, Ho, Wo, C, device=dev, dtype=torch.bfloat16)
    sums = torch.zeros(B, C, device=dev, dtype=torch.int64)
    t_fused = timeit(lambda: cabi.dwconv_bias_silu_nhwc(buf, wt, bias, y, K, S, sums))
    tot += t_fused
    nbytes = (buf.numel() + y.numel()) * 2
    if FAST:
        print("dw B=%d %3dx%3dx%4d k%d s%d: fused %.3f ms (%.0f GB/s)" % (B, H, W, C, K, S, t_fused, nbytes / t_fused / 1e6))
        continue
    xin = buf.permute(0, 3, 1, 2)
    def cudnn_path():
        d = F.conv2d(xin, w, None, stride=S, groups=C).permute(0, 2, 3, 1)
        cabi.bias_silu_nhwc(d.contiguous(), bias, y, sums)
    t_cudnn = timeit(cudnn_path)
    nbytes = (buf.numel() + y.numel()) * 2
    print("dw B=%d %dx%dx%d k%d s%d: fused %.3f ms (%.0f GB/s)   cuDNN+bias_silu %.3f ms" % (B, H, W, C, K, S, t_fused, nbytes / t_fused / 1e6, t_cudnn))
print("total fused %.3f ms" % tot)
