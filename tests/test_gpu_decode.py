"""End to end through the reference-facing API (mp2v_decoder_c via the C ABI): elementary stream in,
YUV out, against the golden hashes of the real reference and -- when it travelled -- the reference
itself (oracle/_ref, serial driver)."""
import numpy as np
import pytest

import oracle_lib as O
from helpers import GOLDEN, GOLDEN_CASES, sha
from tiny_mp2v_dec_b200.decoder import Decoder, frame_bytes
from tiny_mp2v_dec_b200.recon import ReconError
from tiny_mp2v_dec_b200.streamgen import Stream

pytestmark = pytest.mark.gpu


@pytest.fixture(params=[True, False], ids=["device_parser", "host_parser"])
def gpu_vlc(request):
    """both slice parsers behind mp2v_decoder_c: the CUDA one (default) and the host one"""
    return request.param


@pytest.mark.parametrize("name", sorted(GOLDEN_CASES))
def test_decoder_matches_reference_golden(name, gpu_vlc):
    w, h, cf, kw = GOLDEN_CASES[name]
    s = Stream(w, h, cf, **kw)
    d = Decoder(w, h, cf, num_threads=4, gpu_vlc=gpu_vlc)
    yuv = d.decode(s.padded, s.size)
    assert d.stats.frames == GOLDEN[name]["frames"]
    assert sha(yuv) == GOLDEN[name]["yuv_sha256"]
    assert d.stats.launches >= 1 and d.stats.pictures == GOLDEN[name]["frames"]
    # the parser that was asked for is the one that ran
    assert (0 < d.stats.vlc_launches <= GOLDEN[name]["frames"] and d.stats.parse_cpu_seconds == 0.0) if gpu_vlc else (d.stats.vlc_launches == 0 and d.stats.parse_cpu_seconds > 0.0)


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref (the compiled reference) did not travel")
@pytest.mark.parametrize("cf", [1, 2, 3])
def test_decoder_matches_live_reference(cf, gpu_vlc):
    s = Stream(320, 192, cf, seed=50 + cf, n_gops=3, gop_n=10, gop_m=3, qscale_code_max=31, pct_big_levels=10)
    assert Decoder(320, 192, cf, num_threads=3, gpu_vlc=gpu_vlc).decode(s.padded, s.size) == O.ref_decode_serial(s)


@pytest.mark.parametrize("threads,batch,lag", [(1, 1, 1), (2, 4, 2), (8, 8, 6), (16, 32, 12)])
def test_threads_and_batching_do_not_change_the_output(threads, batch, lag, gpu_vlc):
    s = Stream(352, 288, 1, seed=60, n_gops=4, gop_n=12, gop_m=3)
    want = O.oracle_decode_stream(s)
    got = Decoder(352, 288, 1, num_threads=threads, max_batch=batch, output_lag=lag, gpu_vlc=gpu_vlc).decode(s.padded, s.size)
    assert got == want


@pytest.mark.parametrize("cf", [1, 2, 3])
def test_matrices_loaded_once_per_gop_persist(cf, gpu_vlc):
    """quant_matrix_extension (chroma matrices included) only in the first picture of each GOP"""
    s = Stream(176, 144, cf, seed=70 + cf, n_gops=2, gop_n=7, gop_m=3, matrices_once=1)
    assert Decoder(176, 144, cf, num_threads=2, gpu_vlc=gpu_vlc).decode(s.padded, s.size) == O.oracle_decode_stream(s)


@pytest.mark.parametrize("cf,kw", [(1, dict(pct_big_levels=10)), (2, dict(mode=2, pct_intra_in_pb=20)), (3, dict(intra_only=1, pct_field_dct=50))])
def test_intra_vlc_format_0(cf, kw, gpu_vlc):
    """intra blocks coded with table B.14 (intra_vlc_format = 0): outside the reference's envelope (it crashes);
    the oracle reconstructs the generator's ground-truth records"""
    s = Stream(320, 192, cf, seed=80 + cf, n_gops=2, gop_n=5, gop_m=2, intra_vlc_table0=1, **kw)
    assert Decoder(320, 192, cf, num_threads=2, gpu_vlc=gpu_vlc).decode(s.padded, s.size) == O.oracle_decode_stream(s)


def test_no_reordering_gives_coded_order(gpu_vlc):
    s = Stream(176, 144, 1, seed=61, gop_n=7, gop_m=3)
    got = Decoder(176, 144, 1, num_threads=2, reordering=False, gpu_vlc=gpu_vlc).decode(s.padded, s.size)
    disp = O.oracle_decode_stream(s)
    fb = frame_bytes(176, 144, 1)
    frames = {idx: disp[k * fb:(k + 1) * fb] for k, idx in enumerate(s.display_order())}
    assert got == b"".join(frames[i] for i in range(len(s.pictures)))


def test_4k444_decodes(gpu_vlc):
    s = Stream(3840, 2160, 3, seed=62, gop_n=4, gop_m=3, mode=1)
    got = Decoder(3840, 2160, 3, num_threads=8, gpu_vlc=gpu_vlc).decode(s.padded, s.size)
    assert sha(got) == sha(O.oracle_decode_stream(s))


def test_decode_without_download_still_reconstructs(gpu_vlc):
    s = Stream(352, 288, 1, seed=63, gop_n=9, gop_m=3)
    d = Decoder(352, 288, 1, num_threads=2, gpu_vlc=gpu_vlc)
    assert d.decode(s.padded, s.size, want_output=False, download=False) is None
    # nothing comes back but, with the device parser, its 16-byte status per slice and the list of start-code offsets
    assert d.stats.frames == 9 and d.stats.pictures == 9
    assert (9 * 18 * 16 <= d.stats.d2h_bytes < 9 * 18 * 16 + 4096) if gpu_vlc else d.stats.d2h_bytes == 0


def test_malformed_stream_is_an_error_not_a_crash(gpu_vlc):
    from tiny_mp2v_dec_b200.recon import ReconError
    s = Stream(176, 144, 1, seed=64, gop_n=4, gop_m=3)
    bad = s.padded.copy()
    start = int(np.nonzero((bad[:-3] == 0) & (bad[1:-2] == 0) & (bad[2:-1] == 1) & (bad[3:] == 2))[0][1])
    bad[start + 6:start + 60] = 0xFF
    with pytest.raises(ReconError):
        Decoder(176, 144, 1, num_threads=2, gpu_vlc=gpu_vlc).decode(bad, s.size)
    # the decoder stays usable afterwards
    assert sha(Decoder(176, 144, 1, num_threads=2, gpu_vlc=gpu_vlc).decode(s.padded, s.size)) == sha(O.oracle_decode_stream(s))


def test_motion_vector_outside_the_frame_is_rejected():
    """the reference reads out of bounds here (no clamping); the C ABI validates instead"""
    from tiny_mp2v_dec_b200.recon import Recon, ReconError
    s = Stream(64, 48, 1, seed=65, gop_n=2, gop_m=1)
    p = s.pictures[1]
    mb = p.mb.copy()
    k = int(np.nonzero(mb["bits"] & (1 << 30))[0][0])
    mb["mv"][k, 0, 0] = -400
    with Recon(64, 48, 1, n_frames=2, n_pictures=2) as r:
        h0 = r.acquire()
        r.fill(h0, s.pictures[0].params, s.pictures[0].mb, s.pictures[0].coef, dst=0)
        r.submit(h0)
        h1 = r.acquire()
        r.fill(h1, p.params, mb, p.coef, dst=1, l0=0)
        with pytest.raises(ReconError):
            r.submit(h1)


def test_reference_sample_recompiled_against_this_library():
    """drop-in proof: the reference's unmodified sample source (tiny_decoder/tiny_mp2v_dec.cpp), compiled
    against include/core/decoder.h and linked with libmp2v_b200.so by oracle/Makefile, writes the same
    YUV as the reference's own build of it (golden hd422_ipb: hard-wired 1920x1088 4:2:2)."""
    import os
    import subprocess
    import tempfile
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "tiny_mp2v_dec_sample_b200")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/tiny_mp2v_dec_sample_b200 was not built (needs the reference sources at build time)")
    w, h, cf, kw = GOLDEN_CASES["hd422_ipb"]
    s = Stream(w, h, cf, **kw)
    with tempfile.TemporaryDirectory() as d:
        m2v, yuv = os.path.join(d, "in.m2v"), os.path.join(d, "out.yuv")
        s.padded[:s.size + 64].tofile(m2v)      # the sample pads to 16 bytes only; keep the zero tail in the file
        subprocess.check_call([exe, "-v", m2v, "-o", yuv], stdout=subprocess.DEVNULL, timeout=300)
        got = open(yuv, "rb").read()
    assert sha(got) == GOLDEN["hd422_ipb"]["sample_yuv_sha256"]


def test_gop_sharding_over_two_devices(gpu_vlc):
    """closed GOPs dealt round-robin to one pipeline per GPU, frames stitched back in display order;
    no collective, no peer traffic"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    s = Stream(352, 288, 1, seed=70, n_gops=5, gop_n=9, gop_m=3)
    want = O.oracle_decode_stream(s)
    got = Decoder(352, 288, 1, num_threads=4, devices=(0, 1), gpu_vlc=gpu_vlc).decode(s.padded, s.size)
    assert got == want
    one = Decoder(352, 288, 1, num_threads=4, devices=(1,), gpu_vlc=gpu_vlc).decode(s.padded, s.size)
    assert one == want


@pytest.mark.parametrize("n_pipes,w,h,cf,kw", [
    (2, 352, 288, 1, dict(seed=71, n_gops=5, gop_n=9, gop_m=3)),
    (3, 176, 144, 2, dict(seed=72, n_gops=7, gop_n=4, gop_m=2, user_data_bytes=29)),
    (4, 64, 48, 1, dict(seed=73, n_gops=2, gop_n=3, gop_m=1)),              # a stream shorter than the parts' 4 KiB granularity: empty parts
    (8, 640, 368, 1, dict(seed=74, mode=2, n_gops=9, gop_n=6, gop_m=3)),
])
def test_gop_sharding_pipelines_share_the_scan(n_pipes, w, h, cf, kw):
    """several pipelines (here: contexts on ONE GPU, so that a 1-GPU box runs it) each copy and scan one part of the
    stream -- start codes straddle the part boundaries -- and then receive the byte ranges of their own GOPs"""
    s = Stream(w, h, cf, **kw)
    want = O.oracle_decode_stream(s)
    d = Decoder(w, h, cf, num_threads=2, devices=(0,) * n_pipes)
    assert d.decode(s.padded, s.size) == want
    assert d.stats.vlc_launches > 0
    assert d.decode(s.padded, s.size, download=False, want_output=False) is None and d.stats.frames == len(s.pictures)


def test_empty_stream_decodes_to_nothing(gpu_vlc):
    d = Decoder(64, 48, 1, num_threads=2, gpu_vlc=gpu_vlc)
    assert d.decode(np.zeros(512, np.uint8), 0) == b""
    assert d.stats.frames == 0 and d.stats.launches == 0
    s = Stream(64, 48, 1, seed=66, gop_n=1, gop_m=1)     # a single I picture right after
    assert d.decode(s.padded, s.size) == O.oracle_decode_stream(s)


def test_decoder_is_reusable_and_deterministic(gpu_vlc):
    s1 = Stream(176, 144, 2, seed=67, gop_n=6, gop_m=3)
    s2 = Stream(176, 144, 2, seed=68, n_gops=2, gop_n=5, gop_m=2)
    d = Decoder(176, 144, 2, num_threads=3, gpu_vlc=gpu_vlc)
    a1, a2, a3 = d.decode(s1.padded, s1.size), d.decode(s2.padded, s2.size), d.decode(s1.padded, s1.size)
    assert a1 == a3 == O.oracle_decode_stream(s1)
    assert a2 == O.oracle_decode_stream(s2)


def test_bench_workload_full_size_consistency(gpu_vlc):
    """BASELINE.json full size (1080p, the bench's own stream parameters, 2 GOPs of each workload): the two
    product paths -- records reconstructed resident on the device, and the whole decoder from the
    elementary stream -- must agree frame for frame (a checksum of per-frame checksums), and both must equal
    the oracle on the first GOP."""
    import hashlib
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    from tiny_mp2v_dec_b200.recon import reconstruct_stream
    for name in ("1080p420_intra", "1080p420_ipb"):
        wl = bench.WORKLOADS[name]
        g = dict(wl["gen"], n_gops=2)
        s = Stream(wl["width"], wl["height"], wl["chroma_format"], seed=wl["config_id"], **g)
        fb = frame_bytes(wl["width"], wl["height"], wl["chroma_format"])
        a = Decoder(wl["width"], wl["height"], wl["chroma_format"], num_threads=8, gpu_vlc=gpu_vlc).decode(s.padded, s.size)
        b = reconstruct_stream(s)
        sums = lambda y: hashlib.sha256(b"".join(hashlib.sha256(y[i:i + fb]).digest() for i in range(0, len(y), fb))).hexdigest()
        assert len(a) == len(b) == fb * len(s.pictures)
        assert sums(a) == sums(b)
        first = Stream(wl["width"], wl["height"], wl["chroma_format"], seed=wl["config_id"], **dict(g, n_gops=1))
        want = O.oracle_decode_stream(first)
        assert a[:len(want)] == want


def test_concurrent_decoders_share_one_gpu(gpu_vlc):
    """BASELINE.json config 5 in miniature: several independent streams decoded at the same time by separate
    decoder objects on one device (one thread each), every output bit-exact"""
    import threading
    streams = [Stream(352, 288, 1 + (k % 3), seed=400 + k, n_gops=2, gop_n=9, gop_m=3, mode=k % 2) for k in range(6)]
    want = [O.oracle_decode_stream(s) for s in streams]
    got = [None] * len(streams)

    def work(k):
        s = streams[k]
        d = Decoder(352, 288, s.chroma_format, num_threads=2, gpu_vlc=gpu_vlc)
        for _ in range(3):
            got[k] = d.decode(s.padded, s.size)

    threads = [threading.Thread(target=work, args=(k,)) for k in range(len(streams))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert got == want


def test_decode_resident_repeats_the_decode_without_an_upload():
    """mp2v_decoder_c::decode_resident: the stream of the last decode() is still on the device; decoding it again gives
    the same frames, copies no stream bytes and reports the device time of the call"""
    s = Stream(352, 288, 1, seed=77, n_gops=3, gop_n=9, gop_m=3)
    want = O.oracle_decode_stream(s)
    d = Decoder(352, 288, 1, num_threads=2)
    assert d.decode(s.padded, s.size) == want
    first_h2d = d.stats.h2d_bytes
    assert first_h2d >= s.size
    assert d.decode_resident(want_output=True) == want
    assert d.stats.h2d_bytes < first_h2d - s.size // 2 and d.stats.device_ms > 0 and d.stats.vlc_launches > 0
    # the host-parser mode leaves nothing resident
    h = Decoder(352, 288, 1, num_threads=2, gpu_vlc=False)
    assert h.decode(s.padded, s.size) == want
    with pytest.raises(ReconError, match="resident"):
        h.decode_resident()
