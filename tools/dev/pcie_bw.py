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

# D2H of frame-sized chunks while small H2D copies run on another stream (the decoder's traffic pattern)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
hh = torch.empty(64 << 20, dtype=torch.uint8).pin_memory()
dd = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
chunk, small = 3133440, 225 * 1024
for with_h2d in (False, True):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        k = 0
        for o in range(0, n - chunk + 1, chunk):
            with torch.cuda.stream(s1):
                h[o:o + chunk].copy_(d[o:o + chunk], non_blocking=True)
            if with_h2d:
                with torch.cuda.stream(s2):
                    so = (k * small) % ((64 << 20) - small)
                    dd[so:so + small].copy_(hh[so:so + small], non_blocking=True)
                k += 1
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("D2H 3.1 MB chunks%s: %.1f GB/s" % (" + concurrent 225 KB H2D per chunk" if with_h2d else "", reps * (n // chunk) * chunk / dt / 1e9))

# the decoder's pattern: frame-sized D2H into many separate pinned buffers, with and without kernels running
bufs = [torch.empty(chunk, dtype=torch.uint8).pin_memory() for _ in range(58)]
srcs = d[:58 * chunk].view(58, chunk)
x = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
for busy in (False, True):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 8
    for _ in range(reps):
        if busy:
            with torch.cuda.stream(s2):
                for _ in range(6):
                    x = (x @ x).clamp_(-1, 1)
        with torch.cuda.stream(s1):
            for k in range(58):
                bufs[k].copy_(srcs[k], non_blocking=True)
    s1.synchronize()
    dt = time.perf_counter() - t0
    torch.cuda.synchronize()
    print("D2H 58 separate pinned frame buffers%s: %.1f GB/s" % (", GEMMs running" if busy else "", reps * 58 * chunk / dt / 1e9))
