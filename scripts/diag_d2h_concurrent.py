"""Raw pinned D2H bandwidth with one process per GPU copying at the same time (what bounds e2e at N > 1)."""
import os, sys, time, torch, torch.multiprocessing as mp

def work(rank, n_gpus, barrier, q):
    torch.cuda.set_device(rank)
    n = 841 * 1024 * 1024 // 8
    d = torch.empty(n, dtype=torch.float64, device="cuda")
    h = torch.empty(n, dtype=torch.float64, pin_memory=True)
    h.copy_(d, non_blocking=True); torch.cuda.synchronize()
    res = []
    for mode in ("alone", "together"):
        for rep in range(4):
            if mode == "together":
                barrier.wait()
            elif rank != 0:
                continue
            t = time.perf_counter(); h.copy_(d, non_blocking=True); torch.cuda.synchronize()
            res.append((mode, (time.perf_counter() - t) * 1e3))
        barrier.wait()
    q.put((rank, res))

if __name__ == "__main__":
    n = torch.cuda.device_count()
    mp.set_start_method("spawn")
    barrier, q = mp.Barrier(n), mp.Queue()
    ps = [mp.Process(target=work, args=(r, n, barrier, q)) for r in range(n)]
    [p.start() for p in ps]
    out = sorted(q.get() for _ in ps)
    [p.join() for p in ps]
    for rank, res in out:
        print("gpu", rank, " ".join(f"{m}:{ms:.1f}ms" for m, ms in res))
