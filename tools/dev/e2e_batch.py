"""Dev-time: e2e decode throughput of the bench workloads vs launch batching / output lag (device parser)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from tiny_mp2v_dec_b200.decoder import Decoder

for name in sys.argv[1:] or ["1080p420_intra", "1080p420_ipb"]:
    wl = bench.WORKLOADS[name]
    for gops in (8, 24):
        s = bench.make_stream(dict(wl, gen=dict(wl["gen"], n_gops=gops)), 0)
        n = len(s.pictures)
        for batch, lag in [(1, 1), (2, 1), (4, 2), (8, 6), (16, 8)]:
            d = Decoder(wl["width"], wl["height"], wl["chroma_format"], num_threads=14, max_batch=batch, output_lag=lag).prepare(download=True)
            d.decode(s.padded, s.size, want_output=False)
            best = 1e9
            for _ in range(4):
                t0 = time.perf_counter()
                d.decode(s.padded, s.size, want_output=False)
                best = min(best, time.perf_counter() - t0)
            print("%s n=%d batch=%d lag=%d: %.0f fps (%.1f ms)  launches %d" % (name, n, batch, lag, n / best, best * 1e3, d.stats.launches), flush=True)
            d.close()
