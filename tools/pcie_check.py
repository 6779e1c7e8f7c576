"""Pinned host <-> device copy bandwidth of this box (what bounds the e2e figure at one GPU)."""
import torch, time
n = 64 << 20
h = torch.empty(n, dtype=torch.uint8).pin_memory()
h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(both):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10):
        with torch.cuda.stream(s1):
            d.copy_(h, non_blocking=True)
        if both:
            with torch.cuda.stream(s2):
                h2.copy_(d2, non_blocking=True)
    torch.cuda.synchronize(); return n * 10 / (time.perf_counter() - t0) / 1e9
run(False)
print(f"H2D alone {run(False):.1f} GB/s; H2D with concurrent D2H {run(True):.1f} GB/s per direction")
