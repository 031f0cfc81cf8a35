"""Dev-time: per-picture host timeline of one decode() of a bench workload (MP2V_PROFILE=2)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
os.environ["MP2V_TRACE"] = os.environ.get("MP2V_TRACE", "1")
import bench
from tiny_mp2v_dec_b200.decoder import Decoder

name = sys.argv[1] if len(sys.argv) > 1 else "1080p420_intra"
wl = bench.WORKLOADS[name]
s = bench.make_stream(wl, 0)
d = Decoder(wl["width"], wl["height"], wl["chroma_format"], num_threads=14, max_batch=8, output_lag=6).prepare(download=True)
for _ in range(3):
    d.decode(s.padded, s.size, want_output=False)
os.environ["MP2V_PROFILE"] = "2"
d.decode(s.padded, s.size, want_output=False)
