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
