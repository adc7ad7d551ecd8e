"""PCIe host<->device bandwidth of this box with pinned memory (ceiling for the e2e host-buffer path)."""
import torch, time
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h2 = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
def t(fn, it=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(it): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / it
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
dt = t(lambda: d.copy_(h, non_blocking=True)); print("H2D 1 GiB: %.1f GB/s" % (n / dt / 1e9))
dt = t(lambda: h.copy_(d, non_blocking=True)); print("D2H 1 GiB: %.1f GB/s" % (n / dt / 1e9))
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
dt = t(both); print("H2D + D2H concurrently: %.1f + %.1f GB/s" % (n / dt / 1e9, n / dt / 1e9))
