import torch, time
n = 256 << 20
h = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device="cuda")
h2 = torch.empty(n, dtype=torch.uint8).pin_memory(); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def bw(fn, reps=10):
    fn(); torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return reps * n / (time.perf_counter() - t) / 1e9
print("H2D alone %.1f GB/s" % bw(lambda: d.copy_(h, non_blocking=True)))
print("D2H alone %.1f GB/s" % bw(lambda: h.copy_(d, non_blocking=True)))
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
print("H2D + D2H concurrently: %.1f GB/s each direction" % bw(both))
