"""Output path for GPU consumers: NV12 conversion of reconstructed frames on the device,
checked against the interleave of the planar frame the same context hands out."""
import numpy as np
import pytest

import oracle_lib as O
from tiny_mp2v_dec_b200.recon import Recon, ReconError
from tiny_mp2v_dec_b200.streamgen import Stream

pytestmark = pytest.mark.gpu


def _reconstruct_resident(s, r):
    for idx, pic in enumerate(s.pictures):
        h = r.acquire()
        r.fill(h, pic.params, pic.mb, pic.coef, dst=idx, l0=pic.params.l0_frame, l1=pic.params.l1_frame)
        r.submit(h)


@pytest.mark.parametrize("w,h,pitch_extra", [(176, 144, 0), (352, 288, 64), (1920, 1088, 0)])
def test_nv12_matches_the_planar_frame(w, h, pitch_extra):
    import torch
    s = Stream(w, h, 1, seed=320, gop_n=4, gop_m=3, mode=1)
    want = O.oracle_decode_stream(s)
    fb = w * h * 3 // 2
    with Recon(w, h, 1, n_frames=4, n_pictures=4) as r:
        _reconstruct_resident(s, r)
        pitch = w + pitch_extra
        out = torch.full((4, h * 3 // 2, pitch), 0xAB, dtype=torch.uint8, device="cuda")
        r.convert_nv12(0, out[0].data_ptr(), pitch)              # queued behind the launches, no sync in between
        r.convert_nv12_batch([1, 2, 3], [out[f].data_ptr() for f in (1, 2, 3)], pitch)
        r.sync()
        got = out.cpu().numpy()
        for k, f in enumerate(s.display_order()):
            frame = np.frombuffer(want[k * fb:(k + 1) * fb], np.uint8)
            y = frame[:w * h].reshape(h, w)
            cb = frame[w * h:w * h * 5 // 4].reshape(h // 2, w // 2)
            cr = frame[w * h * 5 // 4:].reshape(h // 2, w // 2)
            assert np.array_equal(got[f, :h, :w], y)
            assert np.array_equal(got[f, h:, 0:w:2], cb) and np.array_equal(got[f, h:, 1:w:2], cr)
            assert np.all(got[f, :, w:] == 0xAB)                  # nothing written past the row width


def test_nv12_argument_checks():
    import torch
    s = Stream(64, 48, 1, seed=321, gop_n=1, gop_m=1)
    buf = torch.zeros(64 * 72 + 64, dtype=torch.uint8, device="cuda")
    with Recon(64, 48, 1, n_frames=2, n_pictures=2) as r:
        with pytest.raises(ReconError, match="never been written"):
            r.convert_nv12(0, buf.data_ptr(), 64)
        _reconstruct_resident(s, r)
        with pytest.raises(ReconError, match="frame id"):
            r.convert_nv12(7, buf.data_ptr(), 64)
        with pytest.raises(ReconError, match="16-byte aligned"):
            r.convert_nv12(0, buf.data_ptr() + 4, 64)
        with pytest.raises(ReconError, match="16-byte aligned"):
            r.convert_nv12(0, buf.data_ptr(), 48)
        r.convert_nv12(0, buf.data_ptr(), 64)
        r.sync()
    with Recon(64, 48, 2, n_frames=2, n_pictures=2) as r2:
        with pytest.raises(ReconError, match="4:2:0"):
            r2.convert_nv12(0, buf.data_ptr(), 64)


def _planes(frame, w, h, cf):
    cw, ch = (w if cf == 3 else w // 2), (h // 2 if cf == 1 else h)
    y = frame[:w * h].reshape(h, w)
    cb = frame[w * h:w * h + cw * ch].reshape(ch, cw)
    cr = frame[w * h + cw * ch:].reshape(ch, cw)
    return y, cb, cr


@pytest.mark.parametrize("w,h,pitch_extra", [(176, 144, 0), (1920, 1088, 32)])
def test_p010_and_uyvy_match_the_planar_frame(w, h, pitch_extra):
    """P010 (4:2:0, 16-bit samples with the decoded byte on top) and UYVY (4:2:2 packed) from the device pool"""
    import torch
    for fmt, cf in (("p010", 1), ("uyvy", 2)):
        s = Stream(w, h, cf, seed=322 + cf, gop_n=4, gop_m=3, mode=1)
        want = O.oracle_decode_stream(s)
        fb = w * h * (3 if cf == 1 else 4) // 2
        rows = h * 3 // 2 if fmt == "p010" else h
        pitch = 2 * w + pitch_extra
        with Recon(w, h, cf, n_frames=4, n_pictures=4) as r:
            _reconstruct_resident(s, r)
            out = torch.full((4, rows, pitch), 0xAB, dtype=torch.uint8, device="cuda")
            r.convert_batch(fmt, [0, 1, 2, 3], [out[f].data_ptr() for f in range(4)], pitch)
            r.sync()
            got = out.cpu().numpy()
        for k, f in enumerate(s.display_order()):
            y, cb, cr = _planes(np.frombuffer(want[k * fb:(k + 1) * fb], np.uint8), w, h, cf)
            g = got[f]
            assert np.all(g[:, 2 * w:] == 0xAB)                   # nothing written past the row bytes
            if fmt == "p010":
                s16 = g[:, :2 * w].reshape(rows, w, 2)
                assert np.all(s16[:, :, 0] == 0)                  # low byte: the two padding bits and six more zeros
                assert np.array_equal(s16[:h, :, 1], y)
                assert np.array_equal(s16[h:, 0::2, 1], cb) and np.array_equal(s16[h:, 1::2, 1], cr)
            else:
                px = g[:, :2 * w].reshape(h, w // 2, 4)          # Cb Y0 Cr Y1
                assert np.array_equal(px[:, :, 0], cb) and np.array_equal(px[:, :, 2], cr)
                assert np.array_equal(px[:, :, 1], y[:, 0::2]) and np.array_equal(px[:, :, 3], y[:, 1::2])
    with Recon(64, 48, 3, n_frames=2, n_pictures=2) as r3:
        with pytest.raises(ReconError, match="4:2:2|4:2:0"):
            r3.convert_batch("uyvy", [0], [0x1000], 128)


def test_decoder_hands_device_frames_to_a_gpu_consumer():
    """mp2v_b200_options_t::device_renderer: every frame in display order as DEVICE planes, nothing downloaded; the
    consumer here converts each to NV12 in its own buffer and reads the planes back to compare with the oracle"""
    import ctypes as C
    import torch
    from tiny_mp2v_dec_b200.decoder import Decoder
    from tiny_mp2v_dec_b200 import recon as R
    w, h = 352, 288
    s = Stream(w, h, 1, seed=325, n_gops=2, gop_n=6, gop_m=3)
    want = O.oracle_decode_stream(s)
    fb = w * h * 3 // 2
    frames = []
    nv12 = torch.zeros((len(s.pictures), h * 3 // 2, w), dtype=torch.uint8, device="cuda")
    L = R.lib()

    def consumer(planes, strides, widths, heights, device, frame_id, recon):
        k = len(frames)
        ids, ptrs = (C.c_int32 * 1)(frame_id), (C.c_void_p * 1)(nv12[k].data_ptr())
        assert L.mp2v_recon_convert_frames(C.c_void_p(recon), 0, ids, ptrs, 1, w) == 0
        assert L.mp2v_recon_wait_frame(C.c_void_p(recon), frame_id) == 0
        frames.append((planes[0], strides[0], widths, heights, device))

    d = Decoder(w, h, 1, num_threads=2)
    d.set_device_renderer(consumer, download=False)
    d.decode(s.padded, s.size, want_output=False, download=False)
    assert d.stats.d2h_bytes < fb                                 # no frame crossed PCIe
    torch.cuda.synchronize()
    assert len(frames) == len(s.pictures) and frames[0][2] == [w, w // 2, w // 2] and frames[0][3] == [h, h // 2, h // 2]
    got = nv12.cpu().numpy()
    for k in range(len(s.pictures)):
        y, cb, cr = _planes(np.frombuffer(want[k * fb:(k + 1) * fb], np.uint8), w, h, 1)
        assert np.array_equal(got[k, :h], y) and np.array_equal(got[k, h:, 0::2], cb) and np.array_equal(got[k, h:, 1::2], cr)
    d.close()
