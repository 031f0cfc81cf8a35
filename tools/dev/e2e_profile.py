"""Dev-time: where the end-to-end decode pipeline spends its time (MP2V_PROFILE=1 prints per-stage totals)."""
import os
import sys
import time

os.environ["MP2V_PROFILE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from tiny_mp2v_dec_b200.decoder import Decoder

for name in sys.argv[1:] or ["1080p420_intra", "1080p420_ipb"]:
    wl = bench.WORKLOADS[name]
    s = bench.make_stream(wl, 0)
    n = len(s.pictures)
    for th in (12, 14, 16):
        d = Decoder(wl["width"], wl["height"], wl["chroma_format"], num_threads=th, max_batch=8, output_lag=6).prepare(download=True)
        d.decode(s.padded, s.size, want_output=False)
        t0 = time.perf_counter()
        d.decode(s.padded, s.size, want_output=False)
        dt = time.perf_counter() - t0
        print("%s threads=%d: %.0f fps (%.1f ms)" % (name, th, n / dt, dt * 1e3), flush=True)
        d.close()
