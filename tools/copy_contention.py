"""Dev: how long does a small H2D copy + tiny kernel + 4-byte read on one stream take while another host thread streams a
large pinned upload on a second stream?  (Explains what bench.py's two-thread e2e leg can and cannot overlap.)"""
import sys, threading, time
import torch

dev = torch.device("cuda:0")
big = torch.empty(1200 << 20, dtype=torch.uint8).pin_memory()
big_d = torch.empty_like(big, device=dev)
small = torch.empty(1 << 20, dtype=torch.uint8).pin_memory()
small_d = torch.empty_like(small, device=dev)
out_h = torch.empty(1, dtype=torch.int32).pin_memory()
sx, sy = torch.cuda.Stream(), torch.cuda.Stream()


rate = [0.0]


def uploader(piece, depth, stop):
    evs = [torch.cuda.Event() for _ in range(max(depth, 1))]
    with torch.cuda.stream(sx):
        n_up, t_up = 0, time.perf_counter()
        while not stop.is_set():
            n_up += 1
            k = 0
            for b0 in range(0, big.numel(), piece):
                if depth and k >= depth:
                    evs[k % depth].synchronize()
                big_d[b0:b0 + piece].copy_(big[b0:b0 + piece], non_blocking=True)
                if depth:
                    evs[k % depth].record(sx)
                k += 1
            sx.synchronize()
        rate[0] = n_up * big.numel() / (time.perf_counter() - t_up) / 1e9


def probe(n=200):
    lat = []
    with torch.cuda.stream(sy):
        for _ in range(n):
            t = time.perf_counter()
            small_d.copy_(small, non_blocking=True)
            s = small_d[:1024].to(torch.int32).sum(dtype=torch.int32)
            out_h.copy_(s.reshape(1), non_blocking=True)
            sy.synchronize()
            lat.append((time.perf_counter() - t) * 1e3)
    lat.sort()
    return lat[len(lat) // 2], lat[int(len(lat) * 0.95)], lat[-1]


probe(50)
print("alone: median %.3f p95 %.3f max %.3f ms" % probe())
if len(sys.argv) > 1 and sys.argv[1] == "prio":
    lo, hi = -1, 0
    try:
        lo, hi = torch.cuda.Stream.priority_range()
    except Exception:
        pass
    sy = torch.cuda.Stream(priority=-1)
    print("probe stream at high priority")
for piece, depth in ((1200 << 20, 0), (8 << 20, 3), (8 << 20, 2), (8 << 20, 1), (16 << 20, 1), (32 << 20, 1), (64 << 20, 1)):
    stop = threading.Event()
    th = threading.Thread(target=uploader, args=(piece, depth, stop))
    th.start()
    time.sleep(0.2)
    r = probe()
    stop.set()
    th.join()
    print("upload pieces %4d MB depth %d: median %.3f p95 %.3f max %.3f ms; upload rate %.1f GB/s" % (piece >> 20, depth, *r, rate[0]))
