"""Latency of one bf16 forward + pose decode at small batch, eager vs CUDA-graph mode (development tool).
usage: python scripts/latency_graph.py [batch ...]"""
import sys, time, torch
sys.path.insert(0, ".")
from ccvpe_b200.models import CVM_VIGOR
from ccvpe_b200.synthetic import fill_deterministic, synthetic_pair

dev = torch.device("cuda:0")
m = CVM_VIGOR(dev, circular_padding=True)
fill_deterministic(m.state_dict(), 7)
m.to(dev).eval().set_precision("bf16")
for B in [int(a) for a in sys.argv[1:]] or [1, 4, 16, 64]:
    grd, sat = (t.to(dev) for t in synthetic_pair(B, (320, 640), seed=3))
    res = []
    for graph in (False, True):
        m.set_cuda_graph(graph)
        with torch.no_grad():
            for _ in range(4):
                m.decode_pose(*m(grd, sat)[1:3])
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            K = 20
            for _ in range(K):
                pose = m.decode_pose(*m(grd, sat)[1:3])
                pose["idx"].cpu()                      # a serving loop reads the answer of every request
            torch.cuda.synchronize()
            res.append((time.perf_counter() - t0) / K * 1e3)
    print("B=%3d: eager %.2f ms / forward, CUDA graph %.2f ms / forward (%.1f -> %.1f pairs/s)" %
          (B, res[0], res[1], B / res[0] * 1e3, B / res[1] * 1e3))
