"""Kernel-time table of one bf16 forward (torch profiler), for deciding where encoder time goes.
Usage: python scripts/profile_encoder.py [batch]"""
import sys, torch
sys.path.insert(0, ".")
from ccvpe_b200.models import CVM_VIGOR
from ccvpe_b200.synthetic import fill_deterministic, synthetic_pair

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda:0")
m = CVM_VIGOR(dev, circular_padding=True)
fill_deterministic(m.state_dict(), 7)
m.to(dev).eval().set_precision("bf16")
grd, sat = synthetic_pair(B, (320, 640), seed=3)
grd, sat = grd.to(dev), sat.to(dev)
with torch.no_grad():
    for _ in range(3):
        m(grd, sat)
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3):
            m(grd, sat)
        torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=90))
