"""Host slice parser (product code, runs on the CPU) against the generator's ground truth: every
macroblock record and every coefficient record must come out exactly as it was encoded."""
import numpy as np
import pytest

from tiny_mp2v_dec_b200.abi import MB_FIELD_DCT, mb_coef_off, mb_ncoef
from tiny_mp2v_dec_b200.decoder import parse_stream
from tiny_mp2v_dec_b200.streamgen import Stream

CASES = [
    (128, 64, 1, dict(seed=1, n_gops=2, gop_n=7, gop_m=3)),
    (128, 64, 2, dict(seed=2, n_gops=2, gop_n=7, gop_m=3)),
    (128, 64, 3, dict(seed=3, n_gops=2, gop_n=7, gop_m=3)),
    (128, 64, 1, dict(seed=11, qscale_code_max=31, pct_big_levels=30)),
    (128, 64, 2, dict(seed=12, qscale_code_max=31, pct_big_levels=30, alternate_scan=1, q_scale_type=1)),
    (128, 64, 3, dict(seed=13, qscale_code_max=31, pct_big_levels=50, intra_dc_precision=3)),
    (320, 240, 1, dict(seed=14, n_gops=2, gop_n=15, gop_m=3, mv_range=60)),
    (352, 288, 2, dict(seed=15, gop_n=12, gop_m=4, pct_skipped=40)),
    (64, 48, 1, dict(seed=16, intra_only=1, gop_n=5)),
    (16, 16, 1, dict(seed=21, gop_n=4, gop_m=2)),
    (720, 576, 1, dict(seed=23, gop_n=6, gop_m=3, mode=1)),
    (1920, 1088, 1, dict(seed=18, gop_n=4, gop_m=3)),
    # quant_matrix_extension (with distinct chroma matrices) only in the first picture of a GOP: the matrices persist
    # until the next sequence header (ISO/IEC 13818-2 6.3.11) -- outside the reference's envelope, which needs the
    # extension in every picture (decoder.cpp:187)
    (176, 144, 2, dict(seed=24, n_gops=2, gop_n=7, gop_m=3, matrices_once=1)),
    (176, 144, 3, dict(seed=25, n_gops=2, gop_n=7, gop_m=3, matrices_once=1, alternate_scan=1)),
    (176, 144, 1, dict(seed=26, n_gops=2, gop_n=7, gop_m=3, matrices_once=1)),
    # intra_vlc_format = 0: intra blocks code their AC coefficients with table B.14 -- the reference takes the non-intra
    # first-coefficient path for them and crashes (mb_decoder.cpp:79-88); the generator's ground truth is the oracle
    (176, 144, 1, dict(seed=27, n_gops=2, gop_n=7, gop_m=3, intra_vlc_table0=1, pct_big_levels=10)),
    (176, 144, 3, dict(seed=28, n_gops=1, gop_n=5, gop_m=2, intra_vlc_table0=1, mode=2, pct_intra_in_pb=20)),
    (320, 192, 2, dict(seed=29, intra_only=1, gop_n=3, intra_vlc_table0=1, pct_field_dct=50)),
]


def check_stream(s, threads):
    pics, wall, cpu, n = parse_stream(s.padded, s.size, s.width, s.height, s.chroma_format, threads=threads)
    assert n == len(s.pictures)
    for i, (got, want) in enumerate(zip(pics, s.pictures)):
        assert got.params.picture_coding_type == want.params.picture_coding_type, i
        assert got.params.alternate_scan == want.params.alternate_scan, i
        assert (got.params.l0_frame, got.params.l1_frame) == (want.params.l0_frame, want.params.l1_frame), i
        assert got.gop == want.gop, i
        nsets = 2 if s.chroma_format == 1 else 4
        assert bytes(got.params.W)[:64 * nsets] == bytes(want.params.W)[:64 * nsets], "W differs in picture %d" % i
        assert np.array_equal(got.mb["bits"], want.mb["bits"]), "mb bits differ in picture %d at %s" % (
            i, np.nonzero(got.mb["bits"] != want.mb["bits"])[0][:5])
        assert np.array_equal(got.mb["mv"], want.mb["mv"]), "motion vectors differ in picture %d" % i
        # coefficient offsets differ (chunked arena vs dense), the records must not
        n_coef = mb_ncoef(got.mb["bits"]).astype(np.int64)
        assert np.array_equal(got.mb["coef_off"] & MB_FIELD_DCT, want.mb["coef_off"] & MB_FIELD_DCT), "dct_type differs in picture %d" % i
        idx_g = np.repeat(mb_coef_off(got.mb["coef_off"]).astype(np.int64), n_coef) + (np.arange(n_coef.sum()) - np.repeat(np.cumsum(n_coef) - n_coef, n_coef))
        idx_w = np.repeat(mb_coef_off(want.mb["coef_off"]).astype(np.int64), n_coef) + (np.arange(n_coef.sum()) - np.repeat(np.cumsum(n_coef) - n_coef, n_coef))
        assert np.array_equal(got.coef[idx_g], want.coef[idx_w]), "coefficient records differ in picture %d" % i


@pytest.mark.parametrize("w,h,cf,kw", CASES)
def test_parser_matches_generator_truth(w, h, cf, kw):
    check_stream(Stream(w, h, cf, **kw), threads=1)


def test_parser_threaded_is_deterministic():
    s = Stream(352, 288, 1, seed=5, n_gops=2, gop_n=9, gop_m=3)
    check_stream(s, threads=4)


def test_parser_rejects_garbage():
    from tiny_mp2v_dec_b200.recon import ReconError
    s = Stream(64, 48, 1, seed=9, gop_n=3, gop_m=1)
    bad = s.padded.copy()
    # corrupt the middle of the slice data of the first picture
    start = int(np.nonzero((bad[:-3] == 0) & (bad[1:-2] == 0) & (bad[2:-1] == 1) & (bad[3:] == 1))[0][0])
    bad[start + 6:start + 40] = 0xFF
    with pytest.raises(ReconError):
        parse_stream(bad, s.size, 64, 48, 1)


def test_empty_and_header_only_streams():
    """ragged inputs: nothing, padding only, headers without pictures -> zero pictures, no error"""
    s = Stream(64, 48, 1, seed=9, gop_n=2, gop_m=1)
    for blob in (np.zeros(0, np.uint8), np.zeros(100, np.uint8), s.padded[:int(np.nonzero(s.padded[3:] == 0x00)[0][0])][:40]):
        padded = np.concatenate([np.asarray(blob, np.uint8), np.zeros(256, np.uint8)])
        pics, _, _, n = parse_stream(padded, len(blob), 64, 48, 1)
        assert n == 0 and pics == []


def test_truncated_stream_is_an_error_or_a_prefix():
    """a stream cut in the middle of a slice must not crash: either the cut slice is rejected or the
    pictures before it come out intact"""
    from tiny_mp2v_dec_b200.recon import ReconError
    s = Stream(176, 144, 1, seed=10, gop_n=4, gop_m=1)
    cut = s.size * 2 // 3
    padded = np.concatenate([s.padded[:cut], np.zeros(256, np.uint8)])
    try:
        pics, _, _, n = parse_stream(padded, cut, 176, 144, 1)
        assert 1 <= n <= 4
    except ReconError:
        pass


def test_tall_picture_uses_slice_vertical_position_extension():
    """vertical_size > 2800: slices carry slice_vertical_position_extension (mp2v_hdr.h:347-348)"""
    check_stream(Stream(32, 2816, 1, seed=12, gop_n=3, gop_m=3), threads=2)
