"""Dev-time: throughput of the NV12 output kernel on resident 1080p frames (HBM-bound byte mover)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import json

import torch

from tiny_mp2v_dec_b200.recon import Recon
from tiny_mp2v_dec_b200.streamgen import Stream

w, h, n = 1920, 1088, 60
s = Stream(w, h, 1, seed=2, n_gops=4, gop_n=15, gop_m=1, intra_only=1, mode=1)
peak = 6543.4
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
with Recon(w, h, 1, n_frames=n, n_pictures=8, max_batch=8) as r:
    for idx, pic in enumerate(s.pictures):
        p = r.acquire()
        r.fill(p, pic.params, pic.mb, pic.coef, dst=idx)
        r.submit(p)
    r.sync()
    out = torch.empty((n, h * 3 // 2, w), dtype=torch.uint8, device="cuda")      # 60 frames in + 60 out = 376 MB: larger than L2
    import ctypes as C
    ids = (C.c_int32 * n)(*range(n))
    ptrs = (C.c_void_p * n)(*[out[f].data_ptr() for f in range(n)])
    for reps in (2, 40):
        r.timer_start()
        for _ in range(reps):
            r.L.mp2v_recon_convert_frames_nv12(r.h, ids, ptrs, n, w)      # 2 launches of <= 32 frames, enqueue only
        ms = r.timer_stop()
    fb = w * h * 3 // 2
    gbs = 2 * fb * n * reps / (ms * 1e-3) / 1e9
    print("NV12 conversion: %.2f us/frame, %.0f frames/s, %.0f GB/s algorithmic (read + write) = %.1f %% of %.0f" % (ms * 1e3 / (n * reps), n * reps / (ms * 1e-3), gbs, 100 * gbs / peak, peak))
