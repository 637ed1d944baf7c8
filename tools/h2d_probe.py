"""Host-to-device bandwidth from pinned memory: one copy stream against two (are both copy engines used?)."""
import time, torch
n = 2 << 30
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(two, rows):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    step = n // rows
    for i in range(rows):
        st = (s2 if (two and i % 2) else s1)
        with torch.cuda.stream(st):
            d[i * step:(i + 1) * step].copy_(h[i * step:(i + 1) * step], non_blocking=True)
    torch.cuda.synchronize(); return n / (time.perf_counter() - t0) / 1e9
for rows in (1, 8, 64, 256):
    print(f"rows {rows}: one stream {max(run(False, rows) for _ in range(3)):.1f} GB/s, two streams {max(run(True, rows) for _ in range(3)):.1f} GB/s", flush=True)
