"""Dev-time: end-to-end decode throughput with device-side slice parsing vs the host parser (run on the GPU box).
MP2V_VLC_LANES (1..32) selects how many slices share a warp."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tiny_mp2v_dec_b200.decoder import Decoder
from tiny_mp2v_dec_b200.streamgen import Stream


def one(s, w, h, cf, reps, dl, **kw):
    n = len(s.pictures)
    d = Decoder(w, h, cf, **kw).prepare(download=dl)
    d.decode(s.padded, s.size, want_output=False, download=dl)
    t0 = time.perf_counter()
    for _ in range(reps):
        d.decode(s.padded, s.size, want_output=False, download=dl)
    dt = (time.perf_counter() - t0) / reps
    st = d.stats
    d.close()
    return n / dt, st


def run(name, w, h, cf, reps=4, **kw):
    s = Stream(w, h, cf, **kw)
    n = len(s.pictures)
    print("== %s: %d pictures, %.0f kB/frame, %.0fk coefs/frame" % (name, n, s.size / n / 1e3, sum(len(p.coef) for p in s.pictures) / n / 1e3), flush=True)
    fps, st = one(s, w, h, cf, reps, True, num_threads=14, max_batch=8, output_lag=6)
    print("   host parser  threads=14 download=1: %6.0f fps  launches %d" % (fps, st.launches), flush=True)
    for lanes in (os.environ.get("LANES", "1,2,4,32")).split(","):
        os.environ["MP2V_VLC_LANES"] = lanes
        for dl, batch, lag in [(True, 8, 6), (False, 8, 6), (True, 16, 12), (True, 32, 24)]:
            fps, st = one(s, w, h, cf, reps, dl, num_threads=4, max_batch=batch, output_lag=lag, gpu_vlc=True)
            print("   device parser lanes=%2s download=%d batch=%2d lag=%2d: %6.0f fps   launches %3d  kernel %.2f ms  h2d %.1f MB d2h %.1f MB"
                  % (lanes, dl, batch, lag, fps, st.launches, st.kernel_ms, st.h2d_bytes / 1e6, st.d2h_bytes / 1e6), flush=True)


if __name__ == "__main__":
    run("1080p420 intra natural", 1920, 1088, 1, seed=2, n_gops=8, gop_n=15, gop_m=1, intra_only=1, mode=1, pct_coded=70)
    run("1080p420 IPB natural", 1920, 1088, 1, seed=3, n_gops=8, gop_n=15, gop_m=3, mode=1, pct_coded=70)
