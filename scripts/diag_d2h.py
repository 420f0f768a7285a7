"""Raw pinned D2H / H2D bandwidth of this box (what bounds the e2e number)."""
import time, torch
dev = torch.device("cuda", 0)
for mb in (64, 256, 841):
    n = mb * 1024 * 1024 // 8
    d = torch.empty(n, dtype=torch.float64, device=dev)
    h = torch.empty(n, dtype=torch.float64, pin_memory=True)
    for name, fn in (("D2H", lambda: h.copy_(d, non_blocking=True)), ("H2D", lambda: d.copy_(h, non_blocking=True))):
        fn(); torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            t = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t)
        print(f"{name} {mb} MB: best {mb / 1024 / min(ts):.1f} GiB/s  ({min(ts) * 1e3:.2f} ms)")
# two streams, two halves
n = 841 * 1024 * 1024 // 8
d = torch.empty(n, dtype=torch.float64, device=dev); h = torch.empty(n, dtype=torch.float64, pin_memory=True)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize()
for _ in range(3):
    t = time.perf_counter()
    with torch.cuda.stream(s1): h[: n // 2].copy_(d[: n // 2], non_blocking=True)
    with torch.cuda.stream(s2): h[n // 2:].copy_(d[n // 2:], non_blocking=True)
    torch.cuda.synchronize()
    print(f"D2H 841 MB on 2 streams: {(time.perf_counter() - t) * 1e3:.2f} ms")
