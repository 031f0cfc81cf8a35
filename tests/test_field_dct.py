"""Field DCT (dct_type = 1) in frame pictures with frame_pred_frame_dct = 0 -- the first step of SURVEY.md 8(f)-4.

Pinned by the real reference for 4:2:0 and 4:2:2 (tests/golden: fielddct420_ipb, fielddct422_altscan,
fielddct420_texture; mb_decoder.cpp:172-189).  4:4:4 is NOT pinnable: the reference starts blocks 10 / 11 two rows
down (`(dct_type ? 1 : 8) * stride` with the stride already doubled, mb_decoder.cpp:194-195), writes a row of the
macroblock below and corrupts the heap at the last macroblock row -- there the oracle follows ISO/IEC 13818-2
6.3.17.1 / 7.6 (the lower right-hand blocks start on frame row 1 like the lower left-hand ones) and the CUDA path
is compared with that oracle."""
import numpy as np
import pytest

import oracle_lib as O
from helpers import oracle_decode_parsed, sha
from tiny_mp2v_dec_b200.abi import MB_FIELD_DCT
from tiny_mp2v_dec_b200.decoder import parse_stream
from tiny_mp2v_dec_b200.streamgen import Stream

CASES = [
    (176, 144, 1, dict(seed=601, n_gops=2, gop_n=7, gop_m=3, pct_field_dct=50)),
    (176, 144, 2, dict(seed=602, n_gops=2, gop_n=7, gop_m=3, pct_field_dct=50, pct_skipped=20)),
    (176, 144, 3, dict(seed=603, n_gops=2, gop_n=7, gop_m=3, pct_field_dct=50)),
    (352, 288, 3, dict(seed=604, mode=2, n_gops=1, gop_n=7, gop_m=3, pct_field_dct=70, pct_intra_in_pb=5)),
    (320, 192, 1, dict(seed=605, mode=1, n_gops=2, gop_n=9, gop_m=3, pct_field_dct=100, q_scale_type=1, alternate_scan=1)),
]


def _field_mbs(pics):
    return sum(int(((p.mb["coef_off"] & MB_FIELD_DCT) != 0).sum()) for p in pics)


@pytest.mark.parametrize("w,h,cf,kw", CASES)
def test_host_parser_reads_dct_type(w, h, cf, kw):
    """the generator's ground-truth records carry dct_type in coef_off bit 31; the host parser must find the same"""
    s = Stream(w, h, cf, **kw)
    assert _field_mbs(s.pictures) > 0
    pics, _, _, n = parse_stream(s.padded, s.size, w, h, cf, threads=2)
    assert n == len(s.pictures)
    assert _field_mbs(pics) == _field_mbs(s.pictures)
    for got, want in zip(pics, s.pictures):
        assert np.array_equal(got.mb["coef_off"] & MB_FIELD_DCT, want.mb["coef_off"] & MB_FIELD_DCT)
        assert np.array_equal(got.mb["bits"], want.mb["bits"])
    assert oracle_decode_parsed(pics, w, h, cf) == O.oracle_decode_stream(s)


def test_field_dct_changes_the_picture():
    """the flag is not decoration: the same records reconstructed as frame DCT give another picture"""
    w, h, cf, kw = CASES[0]
    s = Stream(w, h, cf, **kw)
    want = O.oracle_decode_stream(s)
    pics, _, _, _ = parse_stream(s.padded, s.size, w, h, cf, threads=1)
    for p in pics:
        p.mb["coef_off"] &= ~np.uint32(MB_FIELD_DCT)
    assert oracle_decode_parsed(pics, w, h, cf) != want


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref (the compiled reference) is not present")
@pytest.mark.parametrize("cf,mode", [(1, 0), (1, 2), (2, 0), (2, 2)])
def test_oracle_matches_live_reference_on_field_dct(cf, mode):
    s = Stream(176, 144, cf, seed=610 + cf + 10 * mode, n_gops=2, gop_n=7, gop_m=3, mode=mode, pct_field_dct=50, pct_intra_in_pb=5)
    assert O.ref_decode_serial(s) == O.oracle_decode_stream(s)


def test_field_prediction_is_refused():
    """frame_motion_type other than frame-based (field prediction, dual prime) is outside the envelope: an error, not garbage.
    The first P picture's first macroblock with motion gets its frame_motion_type bits (10) patched to 01."""
    from tiny_mp2v_dec_b200.recon import ReconError
    w, h, cf = 176, 144, 1
    s = Stream(w, h, cf, seed=620, n_gops=1, gop_n=4, gop_m=1, pct_field_dct=50, pct_intra_in_pb=0, pct_skipped=0, pct_coded=100)
    bad = s.padded.copy()
    # second picture (P), first slice: slice header = 5 bits quantiser_scale_code + 1 bit extra = 6 bits, then
    # macroblock_address_increment '1' (1 bit), macroblock_type of a P macroblock with motion: '1' (MC, coded) or
    # '001' (MC, not coded) -- every P macroblock of this stream has a pattern, so '1' -- then frame_motion_type (2 bits)
    codes = np.nonzero((bad[:-3] == 0) & (bad[1:-2] == 0) & (bad[2:-1] == 1))[0]
    pic_starts = [int(c) for c in codes if bad[c + 3] == 0x00]
    second = pic_starts[1]
    sl = next(int(c) for c in codes if c > second and bad[c + 3] == 0x01)
    bit = (sl + 4) * 8 + 6 + 1            # first bit of macroblock_type
    def getbit(k):
        return (int(bad[k >> 3]) >> (7 - (k & 7))) & 1

    def setbit(k, v):
        bad[k >> 3] = (int(bad[k >> 3]) & (0xff ^ (1 << (7 - (k & 7))))) | (v << (7 - (k & 7)))
    if getbit(bit) != 1:
        pytest.skip("the first macroblock of the P picture is not 'MC, coded' with this seed")
    assert (getbit(bit + 1), getbit(bit + 2)) == (1, 0)       # frame_motion_type = 10: frame-based
    setbit(bit + 1, 0)
    setbit(bit + 2, 1)                                         # 01: field-based prediction
    with pytest.raises((ReconError, RuntimeError), match="frame-based prediction|slice"):
        parse_stream(bad, s.size, w, h, cf, threads=1)


# ---------------------------------------------------------------------------------------------------- GPU

@pytest.mark.gpu
@pytest.mark.parametrize("w,h,cf,kw", CASES)
@pytest.mark.parametrize("gpu_vlc", [True, False], ids=["device_parser", "host_parser"])
def test_decoder_field_dct_matches_oracle(w, h, cf, kw, gpu_vlc):
    from tiny_mp2v_dec_b200.decoder import Decoder
    s = Stream(w, h, cf, **kw)
    d = Decoder(w, h, cf, num_threads=3, gpu_vlc=gpu_vlc)
    assert d.decode(s.padded, s.size) == O.oracle_decode_stream(s)
    assert (d.stats.vlc_launches > 0) == gpu_vlc


@pytest.mark.gpu
@pytest.mark.parametrize("cf", [1, 2, 3])
def test_cuda_records_path_field_dct_1080p(cf):
    """records straight into the reconstruction kernel (no parser): every macroblock row / lane mapping at full width"""
    from tiny_mp2v_dec_b200.recon import Recon
    w, h = 1920, 1088
    s = Stream(w, h, cf, seed=630 + cf, n_gops=1, gop_n=4, gop_m=3, mode=1, pct_field_dct=50)
    want = O.oracle_decode_stream(s)
    fb = len(want) // len(s.pictures)
    with Recon(w, h, cf, n_frames=4, n_pictures=4) as r:
        for idx, pic in enumerate(s.pictures):
            hnd = r.acquire()
            r.fill(hnd, pic.params, pic.mb, pic.coef, dst=idx, l0=pic.params.l0_frame, l1=pic.params.l1_frame)
            r.submit(hnd)
        r.sync()
        for k, f in enumerate(s.display_order()):
            assert sha(r.download(f)) == sha(want[k * fb:(k + 1) * fb]), "frame %d" % k


@pytest.mark.gpu
def test_device_parser_refuses_field_prediction():
    from tiny_mp2v_dec_b200.decoder import Decoder
    from tiny_mp2v_dec_b200.recon import ReconError
    w, h, cf = 176, 144, 1
    s = Stream(w, h, cf, seed=620, n_gops=1, gop_n=4, gop_m=1, pct_field_dct=50, pct_intra_in_pb=0, pct_skipped=0, pct_coded=100)
    bad = s.padded.copy()
    codes = np.nonzero((bad[:-3] == 0) & (bad[1:-2] == 0) & (bad[2:-1] == 1))[0]
    second = [int(c) for c in codes if bad[c + 3] == 0x00][1]
    sl = next(int(c) for c in codes if c > second and bad[c + 3] == 0x01)
    bit = (sl + 4) * 8 + 6 + 1
    if (int(bad[bit >> 3]) >> (7 - (bit & 7))) & 1 != 1:
        pytest.skip("the first macroblock of the P picture is not 'MC, coded' with this seed")
    for k, v in ((bit + 1, 0), (bit + 2, 1)):
        bad[k >> 3] = (int(bad[k >> 3]) & (0xff ^ (1 << (7 - (k & 7))))) | (v << (7 - (k & 7)))
    d = Decoder(w, h, cf, num_threads=2, gpu_vlc=True)
    with pytest.raises(ReconError, match="frame-based prediction"):
        d.decode(bad, s.size)
    assert d.decode(s.padded, s.size) == O.oracle_decode_stream(s)      # the decoder object survives
