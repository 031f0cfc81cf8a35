"""Dev-time: a few decode() calls of the 1080p IPB texture workload (for ncu captures of the front-end kernels)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
from tiny_mp2v_dec_b200.decoder import Decoder
from tiny_mp2v_dec_b200.streamgen import Stream

TEX = dict(mode=2, texture_noise=3, pct_intra_in_pb=3, q_scale_type=0, alternate_scan=0, intra_dc_precision=0)
s = Stream(1920, 1088, 1, seed=3003, n_gops=8, gop_n=15, gop_m=3, **TEX)
dl = "--download" in sys.argv
import torch
pinned = torch.empty(len(s.padded), dtype=torch.uint8).pin_memory()
pinned.numpy()[:] = s.padded
d = Decoder(1920, 1088, 1, num_threads=4, max_batch=16, output_lag=16).prepare(download=dl)
import time
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    t0 = time.perf_counter()
    d.decode(pinned.numpy(), s.size, want_output=False, download=dl)
    print("decode %.2f ms" % ((time.perf_counter() - t0) * 1e3), file=sys.stderr)
if "--resident" in sys.argv:
    best = 1e9
    for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
        d.decode_resident()
        best = min(best, d.stats.device_ms)
        print("resident %.3f ms (device)" % d.stats.device_ms, file=sys.stderr)
    print("resident best %.3f ms = %.0f frames/s" % (best, len(s.pictures) / best * 1e3), file=sys.stderr)
print("ok", d.stats.launches, d.stats.vlc_launches)
