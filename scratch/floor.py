"""Practical floor for a 2 x 21.6 MB pass on this GPU: plain copy and empty-kernel launch cost,
timed like bench.py (CUDA-graph replay over 8 rotating buffer sets)."""
import torch
dev = torch.device("cuda:0")
R = 8
n = 256 * 21125
src = [torch.randn(n, device=dev) for _ in range(R)]
dst = [torch.empty(n, device=dev) for _ in range(R)]
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for a, b in zip(src, dst):
        b.copy_(a)
    s.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        for a, b in zip(src, dst):
            b.copy_(a)
    tiny = torch.zeros(32, device=dev)
    g2 = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g2, stream=s):
        for _ in range(R):
            tiny.add_(1.0)
    for name, gr in (("copy 21.6MB->21.6MB", g), ("tiny kernel", g2)):
        for _ in range(5):
            gr.replay()
        s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(200):
            gr.replay()
        e1.record(s)
        s.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (200 * R)
        print("%s: %.2f us per launch (%.0f GB/s for 43.3 MB)" % (name, us, 43.3e6 / us / 1e3))
