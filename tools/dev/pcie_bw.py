"""Dev-time: pinned H2D / D2H bandwidth of the box (bounds the end-to-end decode rate: every frame goes D2H)."""
import time
import torch

n = 256 << 20
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, src, dst in (("H2D", h, d), ("D2H", d, h)):
    for chunk in (n, 3133440):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 5
        for _ in range(reps):
            for o in range(0, n - chunk + 1, chunk):
                dst[o:o + chunk].copy_(src[o:o + chunk], non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        moved = reps * (n // chunk) * chunk
        print("%s chunk %9d B: %.1f GB/s" % (name, chunk, moved / dt / 1e9))
