"""Dev-time: BASELINE.json config 5 on one GPU -- 8 concurrent 720p 4:2:0 streams (64 streams over 8 GPUs are
replicas of this), one decoder object + one thread per stream.  usage: multi_stream.py [n_streams] [download 0/1]"""
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
from tiny_mp2v_dec_b200.decoder import Decoder
from tiny_mp2v_dec_b200.streamgen import Stream

n_streams = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dl = (sys.argv[2] != "0") if len(sys.argv) > 2 else True
streams = [Stream(1280, 720, 1, seed=5000 + k, n_gops=4, gop_n=15, gop_m=3, mode=1, pct_coded=70, natural_mean_coefs=5) for k in range(n_streams)]
decs = [Decoder(1280, 720, 1, num_threads=2, max_batch=8, output_lag=6).prepare(download=dl) for _ in streams]
reps = 4
barrier = threading.Barrier(n_streams + 1)


def work(k):
    decs[k].decode(streams[k].padded, streams[k].size, want_output=False, download=dl)
    barrier.wait()
    for _ in range(reps):
        decs[k].decode(streams[k].padded, streams[k].size, want_output=False, download=dl)


ths = [threading.Thread(target=work, args=(k,)) for k in range(n_streams)]
for t in ths:
    t.start()
barrier.wait()
t0 = time.perf_counter()
for t in ths:
    t.join()
dt = time.perf_counter() - t0
frames = sum(len(s.pictures) for s in streams) * reps
print("%d concurrent 720p streams, download=%d: %.0f frames/s in total (%.0f per stream), %.1f GB/s of frame copies"
      % (n_streams, dl, frames / dt, frames / dt / n_streams, frames * 1280 * 720 * 1.5 / dt / 1e9 if dl else 0.0))
