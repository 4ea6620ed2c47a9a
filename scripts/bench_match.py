"""Micro-benchmark of one matching level through the C ABI (development tool)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ccvpe_b200 import cabi

def main():
    B, C, H, R, stride = map(int, sys.argv[1:6])
    backend = cabi.BACKEND_TCGEN05 if (len(sys.argv) > 6 and sys.argv[6] == "tc") else cabi.BACKEND_SIMT
    dev = torch.device("cuda")
    x = torch.randn(B, H, H, C, device=dev, dtype=torch.bfloat16)
    g = torch.randn(B, C, device=dev)
    scores = torch.empty(B, R, H, H, device=dev)
    mx = torch.empty(B, H, H, device=dev); inv = torch.empty(B, H, H, device=dev)
    scratch = torch.empty(cabi.match_scratch_elems(B, C, R), device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    run = lambda: cabi.match_level(x, g, 0, [i * stride for i in range(R)], (1 << R) - 1, scores=scores, max_out=mx, inv_norm=inv, scratch=scratch, backend=backend)
    for _ in range(3): run()
    torch.cuda.synchronize()
    ms = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    best = min(ms)
    nbytes = x.numel() * 2 + (R + 2) * B * H * H * 4
    print("match %s %s: %.4f ms  %.1f GB/s (algorithmic, incl. G build)" % (sys.argv[1:6], "tcgen05" if backend == 2 else "simt", best, nbytes / best / 1e6))
main()
