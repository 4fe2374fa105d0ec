"""Pinned host -> device bandwidth for a 3.5 MB step input: 1 copy vs the same bytes split over k streams."""
import time, torch
dev = torch.device("cuda:0")
n = 3539968
src = torch.empty(n, dtype=torch.uint8).pin_memory()
dst = torch.empty(n, dtype=torch.uint8, device=dev)
for k in (1, 2, 4):
    streams = [torch.cuda.Stream(dev) for _ in range(k)]
    cuts = [n * i // k // 256 * 256 for i in range(k)] + [n]
    def go():
        for i, s in enumerate(streams):
            with torch.cuda.stream(s):
                dst[cuts[i]:cuts[i + 1]].copy_(src[cuts[i]:cuts[i + 1]], non_blocking=True)
    for _ in range(5):
        go()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 200
    for _ in range(reps):
        go()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    print(f"{k} stream(s): {dt * 1e6:.1f} us per {n / 1e6:.2f} MB = {n / dt / 1e9:.1f} GB/s")
big = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
dbig = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
dbig.copy_(big, non_blocking=True); torch.cuda.synchronize()
t0 = time.perf_counter(); dbig.copy_(big, non_blocking=True); torch.cuda.synchronize()
print(f"256 MB single copy: {(256 << 20) / (time.perf_counter() - t0) / 1e9:.1f} GB/s")
# rotating over 40 different pinned sources (141 MB: not cache-resident on the host), as the bench's e2e loop does
srcs = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(40)]
for s_ in srcs:
    s_.fill_(1)
for k in (1, 2):
    streams = [torch.cuda.Stream(dev) for _ in range(k)]
    cuts = [n * i // k // 256 * 256 for i in range(k)] + [n]
    def go(j):
        for i, s in enumerate(streams):
            with torch.cuda.stream(s):
                dst[cuts[i]:cuts[i + 1]].copy_(srcs[j % 40][cuts[i]:cuts[i + 1]], non_blocking=True)
    for j in range(5):
        go(j)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for j in range(200):
        go(j)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 200
    print(f"rotating 40 sources, {k} stream(s): {dt * 1e6:.1f} us = {n / dt / 1e9:.1f} GB/s")
import os
print("cpus", os.cpu_count(), open("/proc/meminfo").readline().strip())
