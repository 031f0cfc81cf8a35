"""Parity of the CUDA reconstruction path (through the C ABI) against the CPU oracle and the golden
hashes of the real reference, on generated streams' ground-truth records."""
import hashlib

import numpy as np
import pytest

import oracle_lib as O
from tiny_mp2v_dec_b200.streamgen import Stream

pytestmark = pytest.mark.gpu

CASES = [
    # (width, height, cf, generator kwargs)
    (128, 64, 1, dict(seed=1, n_gops=2, gop_n=7, gop_m=3)),
    (128, 64, 2, dict(seed=2, n_gops=2, gop_n=7, gop_m=3)),
    (128, 64, 3, dict(seed=3, n_gops=2, gop_n=7, gop_m=3)),
    (128, 64, 1, dict(seed=11, qscale_code_max=31, pct_big_levels=30)),
    (128, 64, 2, dict(seed=12, qscale_code_max=31, pct_big_levels=30, alternate_scan=1, q_scale_type=1)),
    (128, 64, 3, dict(seed=13, qscale_code_max=31, pct_big_levels=50, intra_dc_precision=3)),
    (320, 240, 1, dict(seed=14, n_gops=2, gop_n=15, gop_m=3, mv_range=60)),
    (352, 288, 2, dict(seed=15, gop_n=12, gop_m=4, pct_skipped=40)),
    (64, 48, 1, dict(seed=16, intra_only=1, gop_n=5)),
    (64, 48, 3, dict(seed=17, gop_m=1, gop_n=6)),
    (16, 16, 1, dict(seed=21, gop_n=4, gop_m=2)),          # a single macroblock
    (48, 16, 3, dict(seed=22, gop_n=4, gop_m=3)),
    (640, 480, 3, dict(seed=20, gop_n=6, gop_m=3, all_blocks_coded=1, pct_coded=100)),
    (720, 576, 1, dict(seed=23, gop_n=6, gop_m=3, mode=1)),
]


def _first_diff(a, b, stream):
    fa, fb = np.frombuffer(a, np.uint8), np.frombuffer(b, np.uint8)
    d = np.nonzero(fa != fb)[0]
    fbytes = O.yuv_frame_bytes(stream.width, stream.height, stream.chroma_format)
    return "%d differing bytes, first at frame %d offset %d (gpu %d oracle %d)" % (
        len(d), d[0] // fbytes, d[0] % fbytes, fa[d[0]], fb[d[0]])


@pytest.mark.parametrize("w,h,cf,kw", CASES)
def test_cuda_matches_oracle(w, h, cf, kw):
    from tiny_mp2v_dec_b200.recon import reconstruct_stream
    s = Stream(w, h, cf, **kw)
    want = O.oracle_decode_stream(s)
    got = reconstruct_stream(s)
    assert len(got) == len(want)
    assert got == want, _first_diff(got, want, s)


@pytest.mark.parametrize("w,h,cf,kw", CASES[:3])
def test_unbatched_equals_batched(w, h, cf, kw):
    from tiny_mp2v_dec_b200.recon import reconstruct_stream
    s = Stream(w, h, cf, **kw)
    assert reconstruct_stream(s, batch=False) == reconstruct_stream(s, batch=True)


@pytest.mark.parametrize("cf,seed", [(1, 18), (2, 19)])
def test_cuda_matches_oracle_1080p(cf, seed):
    from tiny_mp2v_dec_b200.recon import reconstruct_stream
    s = Stream(1920, 1088, cf, seed=seed, gop_n=7, gop_m=3)
    want = O.oracle_decode_stream(s)
    got = reconstruct_stream(s)
    assert got == want, _first_diff(got, want, s)
