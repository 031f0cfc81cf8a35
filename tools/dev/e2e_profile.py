"""Dev-time: end-to-end decode of the 1080p IPB texture workload with the decoder's stage profile (MP2V_PROFILE=1)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
from tiny_mp2v_dec_b200.decoder import Decoder
from tiny_mp2v_dec_b200.streamgen import Stream

TEX = dict(mode=2, texture_noise=3, pct_intra_in_pb=3, q_scale_type=0, alternate_scan=0, intra_dc_precision=0)


def run(name, w, h, cf, reps=5, **kw):
    s = Stream(w, h, cf, **kw)
    n = len(s.pictures)
    import torch
    pinned = torch.empty(len(s.padded), dtype=torch.uint8).pin_memory()
    pinned.numpy()[:] = s.padded
    s.padded = pinned.numpy()
    print("== %s: %d pictures, %.0f kB/frame" % (name, n, s.size / n / 1e3), flush=True)
    for dl, batch, lag, gv in [(True, 8, 6, True), (False, 8, 6, True), (True, 16, 16, True), (False, 16, 16, True), (False, 32, 32, True), (True, 8, 6, False)]:
        d = Decoder(w, h, cf, num_threads=14, max_batch=batch, output_lag=lag, gpu_vlc=gv).prepare(download=dl)
        d.decode(s.padded, s.size, want_output=False, download=dl)
        os.environ.pop("MP2V_PROFILE", None)
        t0 = time.perf_counter()
        for _ in range(reps):
            d.decode(s.padded, s.size, want_output=False, download=dl)
        dt = (time.perf_counter() - t0) / reps
        st = d.stats
        print("   gpu_vlc=%d download=%d batch=%2d lag=%2d: %6.0f fps  %.2f ms/call  launches %3d + %3d parse  kernel %.2f ms  h2d %.1f MB d2h %.1f MB"
              % (gv, dl, batch, lag, n / dt, dt * 1e3, st.launches, st.vlc_launches, st.kernel_ms, st.h2d_bytes / 1e6, st.d2h_bytes / 1e6), flush=True)
        os.environ["MP2V_PROFILE"] = "1"
        d.decode(s.padded, s.size, want_output=False, download=dl)
        os.environ.pop("MP2V_PROFILE", None)
        d.close()


if __name__ == "__main__":
    run("1080p420 IPB texture", 1920, 1088, 1, seed=3003, n_gops=8, gop_n=15, gop_m=3, **TEX)
    if "--all" in sys.argv:
        run("1080p420 intra texture", 1920, 1088, 1, seed=2002, n_gops=8, gop_n=15, gop_m=1, intra_only=1, **TEX)
        run("720p420 IPB texture", 1280, 720, 1, seed=5005, n_gops=8, gop_n=15, gop_m=3, **TEX)
