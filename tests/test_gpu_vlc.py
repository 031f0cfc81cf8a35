"""Device-side slice parsing (mp2v_b200_options_t.gpu_vlc / mp2v_recon_submit_slices), what is specific to
it: agreement with the host parser on every syntax knob, error reporting, the envelope and the fall-back.
(tests/test_gpu_decode.py runs the golden vectors and the API tests through BOTH parsers.)"""
import numpy as np
import pytest

import oracle_lib as O
from helpers import sha
from tiny_mp2v_dec_b200.abi import RECON_DEVICE_VLC, RECON_VALIDATE, PicParams
from tiny_mp2v_dec_b200.decoder import Decoder
from tiny_mp2v_dec_b200.recon import Recon, ReconError
from tiny_mp2v_dec_b200.streamgen import Stream

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cf,mode", [(1, 0), (2, 0), (3, 0), (1, 1), (3, 1)])
def test_device_parser_matches_host_parser_and_oracle(cf, mode):
    s = Stream(352, 288, cf, seed=300 + 10 * cf + mode, n_gops=3, gop_n=12, gop_m=3, mode=mode, qscale_code_max=31,
               pct_big_levels=8, q_scale_type=cf & 1, intra_dc_precision=cf - 1, alternate_scan=mode)
    want = O.oracle_decode_stream(s)
    assert Decoder(352, 288, cf, num_threads=4).decode(s.padded, s.size) == want
    assert Decoder(352, 288, cf, num_threads=4, gpu_vlc=True).decode(s.padded, s.size) == want


def test_device_parser_1080p_and_reuse():
    s = Stream(1920, 1088, 1, seed=301, n_gops=2, gop_n=6, gop_m=3, mode=1, pct_coded=70)
    want = sha(O.oracle_decode_stream(s))
    d = Decoder(1920, 1088, 1, num_threads=4, gpu_vlc=True)
    assert sha(d.decode(s.padded, s.size)) == want
    assert sha(d.decode(s.padded, s.size)) == want     # same handle, second stream


def test_device_parser_reports_slice_errors():
    s = Stream(352, 288, 1, seed=302, gop_n=6, gop_m=3)
    buf = s.padded.copy()
    # overwrite the middle of the stream with a pattern that is no valid macroblock syntax
    mid = s.size // 2
    buf[mid:mid + 64] = 0x00
    buf[mid + 64:mid + 96] = 0xFF
    with pytest.raises(ReconError, match="slice"):
        Decoder(352, 288, 1, num_threads=2, gpu_vlc=True).decode(buf, s.size)
    # the decoder object survives: a good stream decodes afterwards
    d = Decoder(352, 288, 1, num_threads=2, gpu_vlc=True)
    with pytest.raises(ReconError):
        d.decode(buf, s.size)
    assert d.decode(s.padded, s.size) == O.oracle_decode_stream(s)


def test_device_parser_rejects_vectors_outside_the_frame():
    # generator knob: no frame clamp on the vectors -> some leave the frame (the reference would read out of bounds)
    s = Stream(176, 144, 1, seed=303, gop_n=9, gop_m=3, mv_range=40, unclamped_mv=1)
    with pytest.raises(ReconError, match="outside the reference frame"):
        Decoder(176, 144, 1, num_threads=2, gpu_vlc=True).decode(s.padded, s.size)
    with pytest.raises(ReconError):
        Decoder(176, 144, 1, num_threads=2).decode(s.padded, s.size)          # the host path rejects it too


def test_submit_slices_needs_a_vlc_context():
    with Recon(176, 144, 1, n_frames=2, n_pictures=2, flags=RECON_VALIDATE) as r:
        pic = r.acquire()
        with pytest.raises(ReconError, match="MP2V_RECON_DEVICE_VLC"):
            r.submit_slices(pic, PicParams(picture_coding_type=1), np.zeros(64, np.uint8), [(0, 8, 1)], ((1, 1), (1, 1)))
        r.release(pic)


def test_submit_slices_checks_its_arguments():
    with Recon(176, 144, 1, n_frames=2, n_pictures=2, flags=RECON_VALIDATE | RECON_DEVICE_VLC) as r:
        data = np.zeros(256, np.uint8)
        pic = r.acquire()
        assert not pic.contents.coef and pic.contents.coef_capacity == 0      # no host coefficient arena in this mode
        with pytest.raises(ReconError, match="picture_coding_type"):
            r.submit_slices(pic, PicParams(picture_coding_type=4), data, [(0, 8, 1)], ((1, 1), (1, 1)))
        with pytest.raises(ReconError, match="missing reference"):
            r.submit_slices(pic, PicParams(picture_coding_type=2), data, [(0, 8, 1)], ((1, 1), (1, 1)))
        with pytest.raises(ReconError, match="one slice per macroblock row"):
            r.submit_slices(pic, PicParams(picture_coding_type=1), data, [(0, 8, 1), (16, 8, 1)], ((1, 1), (1, 1)))
        with pytest.raises(ReconError, match="row outside"):
            r.submit_slices(pic, PicParams(picture_coding_type=1), data, [(0, 8, 40)], ((1, 1), (1, 1)))
        # a picture without slices reconstructs to blank macroblocks
        r.submit_slices(pic, PicParams(picture_coding_type=1), data, [], ((1, 1), (1, 1)), dst=0)
        r.sync()
        assert set(r.download(0)) == {0}


def _start_codes(buf, size):
    b = buf[:size]
    return np.nonzero((b[:-3] == 0) & (b[1:-2] == 0) & (b[2:-1] == 1))[0]


def test_stream_outside_the_device_envelope_takes_the_host_parser():
    """two slices in one macroblock row (legal MPEG-2, handled by the reference): the device parser's
    one-slice-per-row envelope does not hold, so the decoder parses this stream on the host"""
    s = Stream(352, 288, 1, seed=304, gop_n=6, gop_m=3)
    want = O.oracle_decode_stream(s)
    sc = _start_codes(s.padded, s.size)
    # duplicate one slice of the third picture (decoding a slice twice reconstructs the same pixels twice)
    pics = [int(o) for o in sc if s.padded[o + 3] == 0x00]
    k = int(np.searchsorted(sc, pics[2])) + 3
    while not (1 <= s.padded[sc[k] + 3] <= 0xAF):
        k += 1
    a, b = int(sc[k]), int(sc[k + 1])
    buf = np.concatenate([s.padded[:b], s.padded[a:b], s.padded[b:]])
    d = Decoder(352, 288, 1, num_threads=3)
    assert d.decode(buf, s.size + (b - a)) == want
    assert d.stats.vlc_launches == 0 and d.stats.parse_cpu_seconds > 0.0
    # the same decoder object goes back to the device parser for a stream inside the envelope
    assert d.decode(s.padded, s.size) == want
    assert 0 < d.stats.vlc_launches <= len(s.pictures)      # one parse launch per batch of pictures handed over


@pytest.mark.parametrize("cf", [1, 3])
def test_corrupted_streams_never_hang_or_fault(cf):
    """compute-sanitizer is not available on this pool, so memory safety of the device parser is exercised the
    blunt way: 60 seeded corruptions of slice data (byte flips, zero runs, 0xFF runs) through BOTH parsers.
    Every decode must return -- success or ReconError -- and afterwards the same process must still decode the
    clean stream bit-exactly (a faulted kernel would poison the CUDA context for everything after it)."""
    s = Stream(352, 288, cf, seed=310 + cf, n_gops=2, gop_n=9, gop_m=3, qscale_code_max=31, pct_big_levels=5)
    want = O.oracle_decode_stream(s)
    rng = np.random.default_rng(1234 + cf)
    sc = _start_codes(s.padded, s.size)
    slices = [int(o) for o in sc if 1 <= s.padded[o + 3] <= 0xAF]
    outcomes = {True: [0, 0], False: [0, 0]}
    decoders = {v: Decoder(352, 288, cf, num_threads=2, gpu_vlc=v) for v in (True, False)}   # reused across failures
    for trial in range(60):
        buf = s.padded.copy()
        for _ in range(int(rng.integers(1, 4))):
            at = slices[int(rng.integers(0, len(slices)))] + 4 + int(rng.integers(0, 200))
            at = min(at, s.size - 8)
            kind = int(rng.integers(0, 3))
            if kind == 0:
                buf[at] ^= 1 << int(rng.integers(0, 8))
            elif kind == 1:
                buf[at:at + int(rng.integers(1, 12))] = 0xFF
            else:
                buf[at:at + int(rng.integers(1, 6))] = rng.integers(1, 255, dtype=np.uint8)   # no new start codes
        for gpu_vlc in (True, False):
            try:
                decoders[gpu_vlc].decode(buf, s.size)
                outcomes[gpu_vlc][0] += 1
            except ReconError:
                outcomes[gpu_vlc][1] += 1
    assert sum(outcomes[True]) == sum(outcomes[False]) == 60
    assert outcomes[True][1] > 0 and outcomes[False][1] > 0          # some corruptions are detected as syntax errors
    for gpu_vlc in (True, False):
        assert decoders[gpu_vlc].decode(s.padded, s.size) == want


def _coded_pictures(s):
    """walk the start codes of a generated stream: per coded picture its slices' start-code offsets and the f_codes of
    its picture_coding_extension (everything else the C ABI needs is in the generator's ground truth)"""
    sc = _start_codes(s.padded, s.size)
    pics = []
    for o in (int(x) for x in sc):
        code = int(s.padded[o + 3])
        if code == 0x00:
            pics.append(dict(slices=[], f_code=None))
        elif code == 0xB5 and pics and (int(s.padded[o + 4]) >> 4) == 8:
            b = [int(x) for x in s.padded[o + 4:o + 8]]
            pics[-1]["f_code"] = ((b[0] & 15, b[1] >> 4), (b[1] & 15, b[2] >> 4))
        elif 1 <= code <= 0xAF and pics:
            pics[-1]["slices"].append(o)
    assert len(pics) == len(s.pictures)
    return sc, pics


@pytest.mark.parametrize("cf", [1, 2, 3])
def test_stream_resident_api(cf):
    """mp2v_recon_stream_begin / mp2v_recon_submit_stream_picture through the C ABI: the device-side start-code scan
    equals a numpy scan of the same bytes, and pictures handed over as slice offsets decode bit-exactly"""
    s = Stream(352, 288, cf, seed=330 + cf, n_gops=2, gop_n=9, gop_m=3)
    want = O.oracle_decode_stream(s)
    sc, pics = _coded_pictures(s)
    n = len(s.pictures)
    with Recon(352, 288, cf, n_frames=n, n_pictures=6, max_batch=4, flags=RECON_VALIDATE | RECON_DEVICE_VLC) as r:
        codes = r.stream_begin(s.padded, s.size)
        assert codes.tolist() == [int(x) for x in sc]
        for i, (gp, cp) in enumerate(zip(s.pictures, pics)):
            pic = r.acquire()
            r.submit_stream_picture(pic, gp.params, cp["slices"], cp["f_code"], gp.intra_dc_precision, gp.q_scale_type, 1,
                                    dst=i, l0=gp.params.l0_frame, l1=gp.params.l1_frame)
        r.sync()
        st = r.stats()
        assert 0 < st.vlc_launches < n            # batched: fewer parse launches than pictures
        got = b"".join(r.download(i) for i in s.display_order())
    assert got == want


def test_stream_scan_in_parts_equals_the_whole_scan():
    """mp2v_recon_stream_begin(part, deferred) + stream_codes + stream_add: two contexts each copy and scan one part of a
    stream (the cut placed INSIDE a start code prefix), the concatenated lists equal the scan of the whole stream, and a
    context that only ever received its part plus the ranges of the pictures it decodes reconstructs them bit-exactly"""
    s = Stream(352, 288, 1, seed=350, n_gops=2, gop_n=6, gop_m=3)
    want = O.oracle_decode_stream(s)
    fb = len(want) // len(s.pictures)
    sc, pics = _coded_pictures(s)
    whole = [int(x) for x in sc]
    n = len(s.pictures)
    # a 16-byte aligned cut right behind the first byte(s) of some start code prefix: the code begins in part 0, ends in part 1
    cut = next(((o + 2) & ~15) for o in whole[len(whole) // 2:] if ((o + 2) & ~15) > o)
    assert any(o < cut <= o + 2 for o in whole)
    parts = [(0, cut), (cut, s.size - cut)]
    with Recon(352, 288, 1, n_frames=n, n_pictures=n, flags=RECON_VALIDATE | RECON_DEVICE_VLC) as a, \
         Recon(352, 288, 1, n_frames=n, n_pictures=n, flags=RECON_VALIDATE | RECON_DEVICE_VLC) as b:
        for r, part in ((a, parts[0]), (b, parts[1])):
            assert r.stream_begin(s.padded, s.size, part=part, deferred=True) is None
        got_codes = a.stream_codes().tolist() + b.stream_codes().tolist()
        assert got_codes == whole
        with pytest.raises(ReconError, match="no start-code scan is pending"):
            a.stream_codes()
        # context b decodes the second GOP: it holds part 1 only, so the GOP's bytes that lie in part 0 are added
        gop2 = [i for i, p in enumerate(s.pictures) if i >= n // 2]
        lo = min(pics[i]["slices"][0] for i in gop2)
        b.stream_add([(lo, s.size - lo)])
        with pytest.raises(ReconError, match="outside the stream"):
            b.stream_add([(s.size - 4, 64)])
        for k, i in enumerate(gop2):
            gp, cp = s.pictures[i], pics[i]
            ref = lambda f: -1 if f < 0 else f - n // 2
            pic = b.acquire()
            b.submit_stream_picture(pic, gp.params, cp["slices"], cp["f_code"], gp.intra_dc_precision, gp.q_scale_type, 1,
                                    dst=k, l0=ref(gp.params.l0_frame), l1=ref(gp.params.l1_frame))
        b.sync()
        disp = [f for f in s.display_order() if f >= n // 2]
        for f in disp:
            k = s.display_order().index(f)
            assert b.download(f - n // 2) == want[k * fb:(k + 1) * fb]


def test_stream_resident_api_checks_its_arguments():
    s = Stream(176, 144, 1, seed=340, gop_n=3, gop_m=1)
    sc, pics = _coded_pictures(s)
    with Recon(176, 144, 1, n_frames=3, n_pictures=3, flags=RECON_VALIDATE | RECON_DEVICE_VLC) as r:
        pic = r.acquire()
        with pytest.raises(ReconError, match="no resident stream"):
            r.submit_stream_picture(pic, s.pictures[0].params, pics[0]["slices"], ((15, 15), (15, 15)))
        r.stream_begin(s.padded, s.size)
        with pytest.raises(ReconError, match="not a slice start code"):
            r.submit_stream_picture(pic, s.pictures[0].params, [pics[0]["slices"][0] + 1], ((15, 15), (15, 15)))
        with pytest.raises(ReconError, match="outside the stream"):
            r.submit_stream_picture(pic, s.pictures[0].params, [s.size], ((15, 15), (15, 15)))
        with pytest.raises(ReconError, match="one slice per macroblock row"):
            r.submit_stream_picture(pic, s.pictures[0].params, [pics[0]["slices"][0]] * 2, ((15, 15), (15, 15)))
        r.release(pic)
    with Recon(176, 144, 1, n_frames=3, n_pictures=3, flags=RECON_VALIDATE) as r:
        with pytest.raises(ReconError, match="MP2V_RECON_DEVICE_VLC"):
            r.stream_begin(s.padded, s.size)


def test_staged_slices_api_decodes():
    """the per-picture hand-over (mp2v_recon_submit_slices: coded bytes staged and copied per picture) stays bit-exact"""
    s = Stream(352, 288, 1, seed=345, n_gops=2, gop_n=6, gop_m=3)
    want = O.oracle_decode_stream(s)
    sc, pics = _coded_pictures(s)
    ends = {int(a): int(b) for a, b in zip(sc[:-1], sc[1:])}
    n = len(s.pictures)
    with Recon(352, 288, 1, n_frames=n, n_pictures=4, flags=RECON_VALIDATE | RECON_DEVICE_VLC) as r:
        for i, (gp, cp) in enumerate(zip(s.pictures, pics)):
            pic = r.acquire()
            slices = [(o + 4, ends.get(o, s.size) - o - 4, int(s.padded[o + 3])) for o in cp["slices"]]
            r.submit_slices(pic, gp.params, s.padded, slices, cp["f_code"], gp.intra_dc_precision, gp.q_scale_type, 1,
                            dst=i, l0=gp.params.l0_frame, l1=gp.params.l1_frame)
        r.sync()
        assert r.stats().vlc_launches == n
        got = b"".join(r.download(i) for i in s.display_order())
    assert got == want
