"""PCIe probe: pinned H2D, D2H and both at once (what bounds the end-to-end path)."""
import time
import torch
n = 1 << 30
dev = torch.device("cuda", 0)
h_a = torch.empty(n, dtype=torch.uint8).pin_memory()
h_b = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device=dev)
d_b = torch.empty(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(up, down, reps=5):
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        if up:
            with torch.cuda.stream(s1):
                d_a.copy_(h_a, non_blocking=True)
        if down:
            with torch.cuda.stream(s2):
                h_b.copy_(d_b, non_blocking=True)
    torch.cuda.synchronize()
    return n * reps / (time.perf_counter() - t) / 1e9


for _ in range(2):
    run(True, True, 1)
print("H2D only  %.1f GB/s" % run(True, False))
print("D2H only  %.1f GB/s" % run(False, True))
print("both      %.1f GB/s each direction" % run(True, True))
