"""Is the bf16 forward CPU-launch-bound?  Compares the host time to enqueue K steps with their GPU time, and tries the
same step under a CUDA graph (development tool)."""
import sys, time, torch
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
K = 10
with torch.no_grad():
    for _ in range(3):
        out = m(grd, sat)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(K):
        out = m(grd, sat)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("eager: enqueue %.2f ms/step, total %.2f ms/step" % ((t1 - t0) / K * 1e3, (t2 - t0) / K * 1e3))
    # CUDA graph of the same step
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            out = m(grd, sat)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        gout = m(grd, sat)
    torch.cuda.synchronize()
    ref = m(grd, sat)
    g.replay()
    torch.cuda.synchronize()
    outs_g = gout if isinstance(gout, (tuple, list)) else [gout]
    outs_r = ref if isinstance(ref, (tuple, list)) else [ref]
    for a, b in zip(outs_g, outs_r):
        if torch.is_tensor(a):
            print("graph vs eager max abs diff", float((a.float() - b.float()).abs().max()))
    t0 = time.perf_counter()
    for _ in range(K):
        g.replay()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("graph: total %.2f ms/step" % ((t2 - t0) / K * 1e3))
