"""Dev-time: end-to-end decode throughput vs host threads / download / batching (run on the GPU box)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tiny_mp2v_dec_b200.decoder import Decoder, parse_stream
from tiny_mp2v_dec_b200.streamgen import Stream


def run(name, w, h, cf, reps=4, **kw):
    s = Stream(w, h, cf, **kw)
    n = len(s.pictures)
    print("== %s: %d pictures, %.0f kB/frame, %.0fk coefs/frame" % (name, n, s.size / n / 1e3, sum(len(p.coef) for p in s.pictures) / n / 1e3), flush=True)
    for th in (4, 8, 12, 14, 16):
        _, wall, cpu, _ = parse_stream(s.padded, s.size, w, h, cf, threads=th, want_records=False)
        print("   parse only threads=%2d: %6.0f fps  (%.0f fps/core)" % (th, n / wall, n / cpu), flush=True)
    for th, dl, batch, lag in [(8, True, 8, 6), (12, True, 8, 6), (14, True, 8, 6), (16, True, 8, 6), (16, False, 8, 6), (14, False, 8, 6), (14, True, 16, 12), (14, True, 4, 3), (24, True, 8, 6)]:
        d = Decoder(w, h, cf, num_threads=th, max_batch=batch, output_lag=lag).prepare(download=dl)
        d.decode(s.padded, s.size, want_output=False, download=dl)
        t0 = time.perf_counter()
        for _ in range(reps):
            d.decode(s.padded, s.size, want_output=False, download=dl)
        dt = (time.perf_counter() - t0) / reps
        st = d.stats
        print("   e2e threads=%2d download=%d batch=%2d lag=%2d: %6.0f fps   launches %3d  kernel %.2f ms  parse cpu %.3f s (%.0f fps/core)  h2d %.1f MB d2h %.1f MB"
              % (th, dl, batch, lag, n / dt, st.launches, st.kernel_ms, st.parse_cpu_seconds, n / st.parse_cpu_seconds, st.h2d_bytes / 1e6, st.d2h_bytes / 1e6), flush=True)
        d.close()


if __name__ == "__main__":
    run("1080p420 intra fuzz", 1920, 1088, 1, seed=2, n_gops=4, gop_n=15, gop_m=1, intra_only=1)
    run("1080p420 intra natural", 1920, 1088, 1, seed=2, n_gops=4, gop_n=15, gop_m=1, intra_only=1, mode=1)
    run("1080p420 IPB natural", 1920, 1088, 1, seed=3, n_gops=4, gop_n=15, gop_m=3, mode=1)
