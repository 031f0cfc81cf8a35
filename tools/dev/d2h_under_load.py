"""Dev-time: does GPU work slow the copy engine?  Pinned D2H of frame-sized chunks (torch) alone vs while 8 decoders
reconstruct without downloading anything."""
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import torch

from tiny_mp2v_dec_b200.decoder import Decoder
from tiny_mp2v_dec_b200.streamgen import Stream

chunk, nchunk = 3133440, 64
h = torch.empty(chunk * nchunk, dtype=torch.uint8).pin_memory()
d = torch.empty(chunk * nchunk, dtype=torch.uint8, device="cuda")
st = torch.cuda.Stream()


def d2h_rate(seconds):
    moved, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        with torch.cuda.stream(st):
            for k in range(nchunk):
                h[k * chunk:(k + 1) * chunk].copy_(d[k * chunk:(k + 1) * chunk], non_blocking=True)
        st.synchronize()
        moved += chunk * nchunk
    return moved / (time.perf_counter() - t0) / 1e9


print("D2H alone: %.1f GB/s" % d2h_rate(1.0))
streams = [Stream(1280, 720, 1, seed=5000 + k, n_gops=4, gop_n=15, gop_m=3, mode=1, pct_coded=70, natural_mean_coefs=5) for k in range(8)]
decs = [Decoder(1280, 720, 1, num_threads=2).prepare(download=False) for _ in streams]
stop = False
count = [0] * 8


def work(k):
    while not stop:
        decs[k].decode(streams[k].padded, streams[k].size, want_output=False, download=False)
        count[k] += len(streams[k].pictures)


ths = [threading.Thread(target=work, args=(k,)) for k in range(8)]
for t in ths:
    t.start()
time.sleep(0.5)
c0, t0 = sum(count), time.perf_counter()
rate = d2h_rate(2.0)
fps = (sum(count) - c0) / (time.perf_counter() - t0)
stop = True
for t in ths:
    t.join()
print("D2H while 8 decoders run without download (%.0f frames/s): %.1f GB/s" % (fps, rate))
