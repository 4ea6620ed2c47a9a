"""Micro-benchmark of the tcgen05 1x1 conv + bias + SiLU kernel on the EfficientNet-B0 expand / head shapes of both encoders
(development tool).  usage: python scripts/bench_pointwise.py [batch]"""
import os, sys
import torch
from torch.nn import functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ccvpe_b200 import cabi

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
ONLY = int(sys.argv[2]) if len(sys.argv) > 2 else -1      # profile mode: run just this layer index of the ground list, new kernel only
dev = torch.device("cuda")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
# (divisor of the stem resolution, K, N, pad_lo, pad_hi)
LAYERS = [(1, 16, 96, 0, 1), (2, 24, 144, 1, 1), (2, 24, 144, 1, 2), (4, 40, 240, 2, 2), (4, 40, 240, 0, 1), (8, 80, 480, 1, 1),
          (8, 80, 480, 2, 2), (8, 112, 672, 2, 2), (8, 112, 672, 1, 2), (16, 192, 1152, 2, 2), (16, 192, 1152, 1, 1),
          (16, 320, 1280, 0, 0)]
tot_new = tot_old = 0.0
for name, (H0, W0) in (("ground", (160, 320)), ("aerial", (256, 256))):
    for li, (div, K, N, lo, hi) in enumerate(LAYERS):
        if ONLY >= 0 and (li != ONLY or name != "ground"):
            continue
        H, W = H0 // div, W0 // div
        x = torch.randn(B, H, W, K, device=dev, dtype=torch.bfloat16)
        w = (torch.randn(N, K, device=dev) / K ** 0.5).to(torch.bfloat16)
        bias = torch.randn(N, device=dev)
        out = torch.zeros(B, H + lo + hi, W + lo + hi, N, device=dev, dtype=torch.bfloat16)
        w_nk = cabi.pad_k_blocks(w)
        bb = bias.to(torch.bfloat16)

        def new():
            cabi.pointwise_silu_nhwc(x, w_nk, bias, out, lo, hi)

        def old():
            e = F.linear(x.view(-1, K), w, bb).view(B, H, W, N)
            cabi.bias_silu_nhwc(e, None, out[:, lo:lo + H, lo:lo + W, :], None)

        res = []
        if ONLY >= 0:
            for _ in range(3):
                new()
            torch.cuda.synchronize()
            continue
        for fn in (new, old):
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            ms = []
            for _ in range(5):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record(); torch.cuda.synchronize()
                ms.append(e0.elapsed_time(e1))
            res.append(min(ms))
        byts = (x.numel() + B * H * W * N) * 2
        tot_new += res[0]; tot_old += res[1]
        print("%s %3dx%3d K=%3d N=%4d: tcgen05 %.4f ms %6.0f GB/s | linear+bias_silu %.4f ms" %
              (name, H, W, K, N, res[0], byts / res[0] / 1e6, res[1]))
print("total tcgen05 %.3f ms, linear+bias_silu %.3f ms" % (tot_new, tot_old))
