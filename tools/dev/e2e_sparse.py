"""Dev-time: e2e decode rate vs how much parse work a picture carries (is the frame download rate independent of GPU load?)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tiny_mp2v_dec_b200.decoder import Decoder
from tiny_mp2v_dec_b200.streamgen import Stream

for name, kw in [("prediction only", dict(gop_m=3, pct_coded=0, pct_intra_in_pb=0)), ("IPB mean 2 coefs", dict(gop_m=3, pct_coded=40, natural_mean_coefs=2)),
                 ("IPB mean 5", dict(gop_m=3, pct_coded=70, natural_mean_coefs=5)), ("intra mean 6", dict(gop_m=1, intra_only=1, pct_coded=70, natural_mean_coefs=6))]:
    s = Stream(1920, 1088, 1, seed=3, n_gops=8, gop_n=15, mode=1, **kw)
    n = len(s.pictures)
    for dl in (True, False):
        d = Decoder(1920, 1088, 1, num_threads=14, max_batch=8, output_lag=6).prepare(download=dl)
        d.decode(s.padded, s.size, want_output=False, download=dl)
        best = 1e9
        for _ in range(5):
            t0 = time.perf_counter()
            d.decode(s.padded, s.size, want_output=False, download=dl)
            best = min(best, time.perf_counter() - t0)
        print("%-18s %5.0f kB/frame download=%d: %6.0f fps (%.1f ms)%s" % (name, s.size / n / 1e3, dl, n / best, best * 1e3,
              "  d2h %.1f GB/s" % (d.stats.d2h_bytes / best / 1e9) if dl else ""), flush=True)
        d.close()
