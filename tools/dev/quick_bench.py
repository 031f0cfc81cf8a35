"""Dev-time kernel timing: device-resident pictures, CUDA-event time of the reconstruction launches."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tiny_mp2v_dec_b200.recon import Recon
from tiny_mp2v_dec_b200.streamgen import Stream


def run(name, w, h, cf, reps=20, max_batch=128, **kw):
    t0 = time.time()
    s = Stream(w, h, cf, **kw)
    n = len(s.pictures)
    r = Recon(w, h, cf, n_frames=n, n_pictures=n, max_batch=max_batch, flags=0)
    pics = []
    levels = []
    lvl = {}
    for i, p in enumerate(s.pictures):
        hnd = r.acquire()
        r.fill(hnd, p.params, p.mb, p.coef, dst=i, l0=p.params.l0_frame, l1=p.params.l1_frame)
        r.upload(hnd)
        pics.append(hnd)
        l = 1 + max(lvl.get(p.params.l0_frame, -1), lvl.get(p.params.l1_frame, -1))
        lvl[i] = l
        levels.append(l)
    t1 = time.time()
    r.run_resident(pics, levels)   # warm-up
    r.sync()
    r.stats(reset=True)
    r.set_timing(True)
    for _ in range(reps):
        r.run_resident(pics, levels)
    r.sync()
    st = r.stats()
    ms = st.kernel_ms / reps
    gbs = st.algorithmic_bytes / reps / (ms * 1e-3) / 1e9
    print("%-28s pics=%3d launches/rep=%d kernel %.3f ms/rep  %.0f frames/s  alg %.1f MB/frame  %.0f GB/s (%.1f%% of 6543)  [gen+upload %.1fs]"
          % (name, n, st.launches // reps, ms, n / (ms * 1e-3), st.algorithmic_bytes / reps / n / 1e6, gbs, 100 * gbs / 6543.4, t1 - t0), flush=True)
    r.close()


NAT = dict(mode=1, pct_coded=70)
TEX = dict(mode=2, texture_noise=3, pct_intra_in_pb=3, q_scale_type=0, alternate_scan=0, intra_dc_precision=0)

if __name__ == "__main__":
    if "--intra-only" in sys.argv:
        run("1080p420 intra texture", 1920, 1088, 1, reps=3, seed=2002, intra_only=1, gop_n=15, n_gops=2, gop_m=1, **TEX)
        sys.exit(0)
    if "--ipb-only" in sys.argv:
        run("1080p420 IPB texture", 1920, 1088, 1, reps=3, seed=3003, gop_n=15, gop_m=3, n_gops=4, **TEX)
        sys.exit(0)
    run("1080p420 intra natural", 1920, 1088, 1, seed=2, intra_only=1, gop_n=15, n_gops=2, gop_m=1, natural_mean_coefs=6, **NAT)
    run("1080p420 IPB natural", 1920, 1088, 1, seed=3, gop_n=15, gop_m=3, n_gops=4, natural_mean_coefs=5, **NAT)
    run("1080p420 intra texture", 1920, 1088, 1, seed=2002, intra_only=1, gop_n=15, n_gops=2, gop_m=1, **TEX)
    run("1080p420 IPB texture", 1920, 1088, 1, seed=3003, gop_n=15, gop_m=3, n_gops=4, **TEX)
    run("1080p420 intra fuzz", 1920, 1088, 1, seed=2, intra_only=1, gop_n=15, n_gops=2, gop_m=1)
    run("1080p420 IPB fuzz", 1920, 1088, 1, seed=3, gop_n=15, gop_m=3, n_gops=4)
    if "--all" in sys.argv:
        run("1080p422 IPB natural", 1920, 1088, 2, seed=1, gop_n=15, gop_m=3, n_gops=4, natural_mean_coefs=5, **NAT)
        run("720p420 IPB natural", 1280, 720, 1, seed=5, gop_n=15, gop_m=3, n_gops=8, natural_mean_coefs=5, **NAT)
        run("4k444 IPB natural", 3840, 2160, 3, seed=4, gop_n=15, gop_m=3, n_gops=2, natural_mean_coefs=5, **NAT)
    if "--density" in sys.argv:
        # roofline fraction vs coded-block density (prediction-dominated ... transform-dominated)
        for pc in (0, 10, 30, 70, 100):
            run("1080p420 IPB natural coded=%d%%" % pc, 1920, 1088, 1, seed=3, gop_n=15, gop_m=3, n_gops=8, natural_mean_coefs=5, mode=1,
                pct_coded=pc, pct_intra_in_pb=0 if pc == 0 else 8)
