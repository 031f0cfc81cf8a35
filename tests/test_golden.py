"""CPU suite: the oracle (and the generator, and the host parser) against the committed golden
vectors, which hold what the REAL reference decoder produced (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

import oracle_lib as O
from helpers import GOLDEN, GOLDEN_CASES, HERE, oracle_decode_parsed, sha
from tiny_mp2v_dec_b200.decoder import parse_stream
from tiny_mp2v_dec_b200.streamgen import Stream

SMALL = [k for k, v in GOLDEN_CASES.items() if v[0] * v[1] <= 720 * 480]


@pytest.mark.parametrize("name", sorted(GOLDEN_CASES))
def test_generator_is_deterministic(name):
    w, h, cf, kw = GOLDEN_CASES[name]
    s = Stream(w, h, cf, **kw)
    assert sha(s.data.tobytes()) == GOLDEN[name]["stream_sha256"]
    assert len(s.pictures) == GOLDEN[name]["frames"]


@pytest.mark.parametrize("name", sorted(GOLDEN_CASES))
def test_oracle_matches_reference_golden(name):
    w, h, cf, kw = GOLDEN_CASES[name]
    s = Stream(w, h, cf, **kw)
    assert sha(O.oracle_decode_stream(s)) == GOLDEN[name]["yuv_sha256"]


@pytest.mark.parametrize("name", SMALL)
def test_parser_plus_oracle_matches_reference_golden(name):
    """host parser (product) -> records -> oracle reconstruction == the reference's YUV"""
    w, h, cf, kw = GOLDEN_CASES[name]
    s = Stream(w, h, cf, **kw)
    pics, _, _, _ = parse_stream(s.padded, s.size, w, h, cf, threads=2)
    assert sha(oracle_decode_parsed(pics, w, h, cf)) == GOLDEN[name]["yuv_sha256"]


def test_raw_fixture_without_generator():
    """committed .m2v / .yuv pair (reference output): parser + oracle, no generator involved"""
    m2v = np.fromfile(os.path.join(HERE, "golden", "tiny420_m1.m2v"), np.uint8)
    want = open(os.path.join(HERE, "golden", "tiny420_m1.yuv"), "rb").read()
    padded = np.concatenate([m2v, np.zeros(256, np.uint8)])
    pics, _, _, n = parse_stream(padded, m2v.size, 48, 32, 1)
    assert n == 10
    assert oracle_decode_parsed(pics, 48, 32, 1) == want


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref (the compiled reference) is not present")
@pytest.mark.parametrize("cf", [1, 2, 3])
@pytest.mark.parametrize("seed", [31, 32, 33])
def test_oracle_matches_live_reference(cf, seed):
    """the pin itself, live: random streams through the unmodified reference (serial driver)"""
    s = Stream(160, 96, cf, seed=seed, n_gops=2, gop_n=8, gop_m=3, qscale_code_max=31 if seed == 33 else 12,
               pct_big_levels=25 if seed == 33 else 3)
    assert O.ref_decode_serial(s) == O.oracle_decode_stream(s)


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref (the compiled reference) is not present")
def test_reference_mt_equals_serial_on_large_pictures():
    """the MT scheduler races on tiny pictures; at 720p and above it agrees with serial"""
    s = Stream(1280, 720, 1, seed=44, gop_n=7, gop_m=3)
    n, yuv, _ = O.ref_decode_mt(s, threads=4)
    assert n == len(s.pictures)
    assert yuv == O.ref_decode_serial(s)
