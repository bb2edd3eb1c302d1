"""H2D bandwidth from pinned host memory: one copy vs the same bytes split over two streams, and with a D2H running."""
import torch, time, json
dev = torch.device("cuda:0")
res = {}
for mb in (4, 21.67, 64):
    n = int(mb * 1e6) // 4
    h = torch.empty(n, dtype=torch.float32).pin_memory()
    d = torch.empty(n, dtype=torch.float32, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    def one():
        with torch.cuda.stream(s1):
            d.copy_(h, non_blocking=True)
    def two():
        half = n // 2
        with torch.cuda.stream(s1):
            d[:half].copy_(h[:half], non_blocking=True)
        with torch.cuda.stream(s2):
            d[half:].copy_(h[half:], non_blocking=True)
    def four():
        q = n // 4
        for i, s in enumerate((s1, s2, s1, s2)):
            with torch.cuda.stream(s):
                d[i * q:(i + 1) * q].copy_(h[i * q:(i + 1) * q], non_blocking=True)
    for name, fn in (("one", one), ("two_streams", two), ("four_chunks", four)):
        for _ in range(5): fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 50
        for _ in range(reps): fn()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        res["%s_%.0fMB" % (name, mb)] = round(n * 4 / dt / 1e9, 1)
print(json.dumps(res))
