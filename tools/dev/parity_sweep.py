"""Dev-time: randomized parity sweep -- many small streams with random geometry / syntax knobs through both
slice parsers of the decode API, each compared bit for bit with the oracle.  usage: parity_sweep.py [n] [seed]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import numpy as np

import oracle_lib as O
from tiny_mp2v_dec_b200.decoder import Decoder
from tiny_mp2v_dec_b200.streamgen import Stream

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
GEOMS = [(16, 16), (32, 16), (16, 48), (48, 32), (64, 48), (80, 80), (176, 144), (208, 112), (352, 288), (416, 240), (720, 576), (1280, 720), (64, 2832)]
bad = 0
t0 = time.time()
for k in range(n):
    w, h = GEOMS[int(rng.integers(0, len(GEOMS)))]
    cf = int(rng.integers(1, 4))
    big = w * h > 300000
    kw = dict(seed=int(rng.integers(1, 1 << 30)), n_gops=int(rng.integers(1, 3)), gop_n=int(rng.integers(1, 5 if big else 13)),
              gop_m=int(rng.integers(1, 4)), mode=int(rng.integers(0, 3)), qscale_code_max=int(rng.choice([4, 12, 31])),
              alternate_scan=int(rng.integers(-1, 2)), q_scale_type=int(rng.integers(-1, 2)), intra_dc_precision=int(rng.integers(-1, 4)),
              pct_skipped=int(rng.choice([0, 15, 60])), pct_intra_in_pb=int(rng.choice([0, 10, 50])), pct_coded=int(rng.choice([0, 30, 70, 100])),
              pct_mb_quant=int(rng.choice([0, 10, 50])), pct_big_levels=int(rng.choice([0, 3, 30])), all_blocks_coded=int(rng.integers(0, 2)),
              mv_range=int(rng.choice([0, 3, 24, 100])), pct_field_dct=int(rng.choice([0, 0, 30, 100])), matrices_once=int(rng.integers(0, 2)), intra_vlc_table0=int(rng.integers(0, 2)))
    if rng.integers(0, 4) == 0:
        kw["intra_only"] = 1
    try:
        s = Stream(w, h, cf, **kw)
    except Exception as e:
        print("skip (generator): %s %s" % ((w, h, cf), e))
        continue
    want = O.oracle_decode_stream(s)
    for gpu_vlc in (True, False):
        try:
            got = Decoder(w, h, cf, num_threads=int(rng.integers(1, 6)), max_batch=int(rng.choice([1, 3, 8, 32])), output_lag=int(rng.integers(1, 8)),
                          gpu_vlc=gpu_vlc).decode(s.padded, s.size)
        except Exception as e:
            got = repr(e)
        if got != want:
            bad += 1
            print("MISMATCH gpu_vlc=%s %dx%d cf=%d %s -> %s" % (gpu_vlc, w, h, cf, kw, got if isinstance(got, str) else "different YUV"), flush=True)
print("parity sweep: %d streams x 2 parsers, %d mismatches, %.0f s" % (n, bad, time.time() - t0))
