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
    # IDCT variant selection: small levels (saturation-free pass 2), a mix around the bound, all-exact
    (352, 288, 2, dict(seed=24, gop_n=6, gop_m=3, mode=1, qscale_code_max=4)),
    (352, 288, 1, dict(seed=25, gop_n=6, gop_m=3, pct_big_levels=0, qscale_code_max=6)),
    (352, 288, 3, dict(seed=26, gop_n=6, gop_m=3, pct_big_levels=1, qscale_code_max=20, q_scale_type=1)),
    (352, 288, 1, dict(seed=27, gop_n=4, gop_m=3, pct_big_levels=100, qscale_code_max=31, q_scale_type=1)),
    (1280, 720, 1, dict(seed=28, gop_n=4, gop_m=3, mode=1)),
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


def test_out_of_range_intra_dc_takes_the_exact_path():
    """records-level: intra DC values far outside what a conforming stream can carry (the host ships
    wrap16(pred << shift), any int16 is representable) must still match the reference arithmetic"""
    import ctypes as C
    from tiny_mp2v_dec_b200.abi import COEF_RAW
    from tiny_mp2v_dec_b200.recon import reconstruct_stream
    s = Stream(176, 144, 1, seed=29, intra_only=1, gop_n=2)
    rng = np.random.default_rng(5)
    for p in s.pictures:
        raw = np.nonzero(p.coef & COEF_RAW)[0]
        pick = raw[rng.random(len(raw)) < 0.2]
        vals = rng.integers(-32768, 32768, len(pick)).astype(np.int64)
        p.coef[pick] = (p.coef[pick] & np.uint32(0xffff0000)) | (vals & 0xffff).astype(np.uint32)
    want = O.oracle_decode_stream(s)
    assert reconstruct_stream(s) == want


def test_bound_weights_cover_the_range_analysis():
    """the kernel's c_bound_w table must dominate 16 * Omax[k] * G[c] from tools/dev/idct_bounds.py"""
    import os
    import re
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "tools", "dev"))
    import idct_bounds
    log, outs = idct_bounds.analyse()
    G = np.max(np.array([np.abs(v.c) for _, v in log]), axis=0)
    Omax = np.array([np.abs(o.c) for o in outs]).max(axis=0)
    E = max(v.e for _, v in log)
    src = open(os.path.join(root, "tiny_mp2v_dec_b200", "csrc", "recon_kernels.cu")).read()
    body = re.search(r"c_bound_w\[64\] = \{(.*?)\};", src, re.S).group(1)
    w = np.array([int(x) for x in re.findall(r"\d+", body)]).reshape(8, 8)
    assert (w >= 16 * np.outer(Omax, G)).all()
    limit = int(re.search(r"kBoundLimit = 16 \* (\d+)", src).group(1))
    assert limit <= 32767 - (G.sum() + 1) * E
    # pass 1 (MODE 1): with |F0| <= 3036 and |F| <= 2048 only the 8 outputs can leave int16
    X = np.array([3036.0] + [2048.0] * 7)
    for name, v in log:
        if not name.startswith("o"):
            assert np.abs(v.c) @ X + v.e < 32767, name
