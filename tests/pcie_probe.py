"""Measures pinned host<->device copy rates of the box (context for the e2e number; not a test)."""
import time

import torch

n = 1 << 30
h = torch.empty(n, dtype=torch.float32).pin_memory()
d = torch.empty(n, dtype=torch.float32, device="cuda")
for name, fn in (("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(d, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    print(name, "%.1f GB/s" % (3 * n * 4 / (time.perf_counter() - t) / 1e9))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
h2 = torch.empty(n // 4, dtype=torch.float32).pin_memory()
d2 = torch.empty(n // 4, dtype=torch.float32, device="cuda")
torch.cuda.synchronize()
t = time.perf_counter()
with torch.cuda.stream(s1):
    for _ in range(3):
        d.copy_(h, non_blocking=True)
with torch.cuda.stream(s2):
    for _ in range(12):
        h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t
print("bidirectional: h2d %.1f GB/s + d2h %.1f GB/s" % (3 * n * 4 / dt / 1e9, 12 * (n // 4) * 4 / dt / 1e9))
