"""Dev-time: where the end-to-end decode pipeline spends its time (MP2V_PROFILE=1 prints per-stage totals).
usage: e2e_profile.py [workload ...] ; env DOWNLOAD=0/1, REPS, HOST=1 (host parser too)"""
import os
import sys
import time

os.environ["MP2V_PROFILE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from tiny_mp2v_dec_b200.decoder import Decoder

dl = os.environ.get("DOWNLOAD", "1") != "0"
reps = int(os.environ.get("REPS", "3"))
for name in sys.argv[1:] or ["1080p420_intra", "1080p420_ipb"]:
    wl = bench.WORKLOADS[name]
    s = bench.make_stream(wl, 0)
    n = len(s.pictures)
    modes = [True] + ([False] if os.environ.get("HOST") else [])
    for gpu_vlc in modes:
        d = Decoder(wl["width"], wl["height"], wl["chroma_format"], num_threads=14, max_batch=8, output_lag=6, gpu_vlc=gpu_vlc).prepare(download=dl)
        d.decode(s.padded, s.size, want_output=False, download=dl)
        for _ in range(reps):
            t0 = time.perf_counter()
            d.decode(s.padded, s.size, want_output=False, download=dl)
            dt = time.perf_counter() - t0
            print("%s gpu_vlc=%d download=%d: %.0f fps (%.1f ms)  launches %d + %d" % (name, gpu_vlc, dl, n / dt, dt * 1e3, d.stats.launches, d.stats.vlc_launches), flush=True)
        d.close()
