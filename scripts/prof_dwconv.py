import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ccvpe_b200 import cabi
dev = torch.device("cuda")
B, H, W, C, K, S = 64, 40, 80, 240, 5, 1
lo = hi = 2
buf = torch.randn(B, H + 4, W + 4, C, device=dev).to(torch.bfloat16)
wt = (torch.randn(K * K, C, device=dev) * 0.3).to(torch.bfloat16)
bias = torch.randn(C, device=dev).to(torch.bfloat16)
y = torch.empty(B, H, W, C, device=dev, dtype=torch.bfloat16)
sums = torch.zeros(B, C, device=dev, dtype=torch.int64)
for _ in range(3):
    cabi.dwconv_bias_silu_nhwc(buf, wt, bias, y, K, S, sums)
torch.cuda.synchronize()
