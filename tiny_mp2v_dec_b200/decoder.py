"""Python mirror of the reference-facing decode API (include/mp2v_decoder.hpp via mp2v_decode_c.h).

`Decoder` mirrors mp2v_decoder_c: construct with the fields of decoder_config_t, call decode() once
with a whole elementary stream.  `parse_stream` exposes the host-only slice parser (tests, host-parse
throughput).  No CPU reconstruction exists behind these calls."""
import ctypes as C

import numpy as np

from .abi import OK, MbInfo, PicParams, mb_dtype
from .recon import ReconError, lib as _recon_lib


class DecodeParams(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("chroma_format", C.c_int32),
                ("pictures_pool_size", C.c_int32), ("num_threads", C.c_int32), ("reordering", C.c_int32),
                ("n_devices", C.c_int32), ("devices", C.c_int32 * 8), ("max_batch", C.c_int32), ("output_lag", C.c_int32),
                ("download_frames", C.c_int32), ("hash_output", C.c_int32), ("host_parser", C.c_int32)]


class DecodeStats(C.Structure):
    _fields_ = [("frames", C.c_uint64), ("pictures", C.c_uint64), ("launches", C.c_uint64), ("h2d_bytes", C.c_uint64),
                ("d2h_bytes", C.c_uint64), ("algorithmic_bytes", C.c_uint64), ("kernel_ms", C.c_double),
                ("parse_cpu_seconds", C.c_double), ("wall_seconds", C.c_double), ("hash", C.c_uint64),
                ("vlc_launches", C.c_uint64), ("device_ms", C.c_double)]


# void (*)(void* user, void* const planes[3], const int32 strides[3], const int32 widths[3], const int32 heights[3], int32 device, int32 frame_id, mp2v_recon_t*)
DEVICE_FRAME_FN = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                              C.c_int32, C.c_int32, C.c_void_p)

DECODE_EXPORTS = ["mp2v_decode_stream", "mp2v_decoder_create", "mp2v_decoder_decode", "mp2v_decoder_decode_resident", "mp2v_decoder_set_device_renderer", "mp2v_decoder_destroy", "mp2v_parse_stream", "mp2v_parsed_num_pictures", "mp2v_parsed_picture",
                  "mp2v_parsed_wall_seconds", "mp2v_parsed_cpu_seconds", "mp2v_parsed_free"]

_bound = False


def lib():
    global _bound
    L = _recon_lib()
    if not _bound:
        P = C.POINTER
        L.mp2v_decode_stream.argtypes = [P(DecodeParams), C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                         P(C.c_size_t), P(DecodeStats), C.c_char_p, C.c_size_t]
        L.mp2v_decoder_create.argtypes = [P(DecodeParams), P(C.c_void_p), C.c_char_p, C.c_size_t]
        L.mp2v_decoder_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                          P(C.c_size_t), P(DecodeStats), C.c_char_p, C.c_size_t]
        L.mp2v_decoder_decode_resident.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, P(C.c_size_t), P(DecodeStats),
                                                   C.c_char_p, C.c_size_t]
        L.mp2v_decoder_set_device_renderer.argtypes = [C.c_void_p, DEVICE_FRAME_FN, C.c_void_p]
        L.mp2v_decoder_destroy.argtypes = [C.c_void_p]
        L.mp2v_decoder_destroy.restype = None
        L.mp2v_parse_stream.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, P(C.c_void_p), C.c_char_p, C.c_size_t]
        L.mp2v_parsed_num_pictures.argtypes = [C.c_void_p]
        L.mp2v_parsed_picture.argtypes = [C.c_void_p, C.c_int, P(PicParams), P(P(MbInfo)), P(P(C.c_uint32)), P(C.c_uint32),
                                          P(C.c_int32), P(C.c_int32)]
        L.mp2v_parsed_wall_seconds.argtypes = [C.c_void_p]
        L.mp2v_parsed_wall_seconds.restype = C.c_double
        L.mp2v_parsed_cpu_seconds.argtypes = [C.c_void_p]
        L.mp2v_parsed_cpu_seconds.restype = C.c_double
        L.mp2v_parsed_free.argtypes = [C.c_void_p]
        L.mp2v_parsed_free.restype = None
        _bound = True
    return L


def frame_bytes(width, height, chroma_format):
    return width * height * {1: 3, 2: 4, 3: 6}[chroma_format] // 2


class Decoder:
    """mp2v_decoder_c(decoder_config_t{width, height, chroma_format, pictures_pool_size, num_threads, reordering})"""

    def __init__(self, width, height, chroma_format, pictures_pool_size=10, num_threads=8, reordering=True,
                 devices=(0,), max_batch=0, output_lag=0, gpu_vlc=True):
        """gpu_vlc (mp2v_b200_options_t.gpu_vlc, default on): slices are parsed on the device and the host only finds
        start codes; False forces the host slice parser on num_threads threads"""
        self.p = DecodeParams(width, height, chroma_format, pictures_pool_size, num_threads, 1 if reordering else 0,
                              len(devices), (C.c_int32 * 8)(*devices), max_batch, output_lag, 1, 0, 0 if gpu_vlc else 1)
        self.stats = None
        self.h = None
        self._download = None

    def _handle(self, download):
        """the native decoder (device contexts allocated here, as the reference allocates its frame pool in
        the constructor); re-created when the download mode changes"""
        if self.h is not None and self._download == download:
            return self.h
        self.close()
        L = lib()
        self.p.download_frames = 1 if download else 0
        h = C.c_void_p()
        err = C.create_string_buffer(512)
        rc = L.mp2v_decoder_create(C.byref(self.p), C.byref(h), err, 512)
        if rc != OK:
            raise ReconError("decoder create failed (%d): %s" % (rc, err.value.decode()))
        self.h, self._download = h, download
        return h

    def prepare(self, download=True):
        self._handle(download)
        return self

    def close(self):
        if self.h is not None:
            lib().mp2v_decoder_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def decode(self, padded, size, want_output=True, download=True):
        """padded: uint8 array holding the stream followed by >= 64 bytes of padding; size: stream length.
        Returns cropped planar YUV (display order) as bytes when want_output, else None; self.stats is filled."""
        L = lib()
        buf = np.ascontiguousarray(padded, np.uint8)
        assert buf.size >= size + 64, "decode(): the stream buffer must be padded with >= 64 bytes"
        h = self._handle(download)
        st = DecodeStats()
        err = C.create_string_buffer(512)
        out = None
        cap = 0
        if want_output and download:
            # upper bound: every picture start code yields one frame
            n_pics = 0 if size < 4 else int(np.count_nonzero((buf[:size - 3] == 0) & (buf[1:size - 2] == 0) & (buf[2:size - 1] == 1) & (buf[3:size] == 0)))
            cap = n_pics * frame_bytes(self.p.width, self.p.height, self.p.chroma_format)
            out = np.empty(max(cap, 1), np.uint8)
        self._last_cap = cap        # (0 unless this call produced output: decode_resident(want_output=True) then counts the pictures itself)
        self._keep, self._keep_size = buf, size      # decode_resident needs the bytes to stay alive
        nbytes = C.c_size_t()
        rc = L.mp2v_decoder_decode(h, buf.ctypes.data, size, None, None, out.ctypes.data if out is not None else None,
                                   cap, C.byref(nbytes), C.byref(st), err, 512)
        self.stats = st
        if rc != OK:
            raise ReconError("decode failed (%d): %s" % (rc, err.value.decode()))
        if out is None:
            return None
        assert nbytes.value <= cap, (nbytes.value, cap)
        return out[:nbytes.value].tobytes()


def _decode_resident(self, want_output=False):
    """mp2v_decoder_c::decode_resident: decode once more the stream the last decode() left on the device (no upload, no
    start-code scan); the buffer given to that decode() must still be alive.  Returns the YUV when want_output."""
    L = lib()
    st = DecodeStats()
    err = C.create_string_buffer(512)
    out, cap = None, 0
    if want_output and self._download:
        if not self._last_cap:
            b = self._keep[:self._keep_size]
            n_pics = 0 if len(b) < 4 else int(np.count_nonzero((b[:-3] == 0) & (b[1:-2] == 0) & (b[2:-1] == 1) & (b[3:] == 0)))
            self._last_cap = n_pics * frame_bytes(self.p.width, self.p.height, self.p.chroma_format)
        cap = self._last_cap
        out = np.empty(max(cap, 1), np.uint8)
    nbytes = C.c_size_t()
    rc = L.mp2v_decoder_decode_resident(self.h, None, None, out.ctypes.data if out is not None else None, cap, C.byref(nbytes),
                                        C.byref(st), err, 512)
    self.stats = st
    if rc != OK:
        raise ReconError("decode_resident failed (%d): %s" % (rc, err.value.decode()))
    return out[:nbytes.value].tobytes() if out is not None else None


Decoder.decode_resident = _decode_resident


def _set_device_renderer(self, fn, download=False):
    """fn(planes[3] device pointers, strides[3], widths[3], heights[3], device, frame_id, recon handle) is called on the
    decoder's output thread for every frame in display order (mp2v_b200_options_t::device_renderer); None removes it"""
    h = self._handle(download)
    self._dev_cb = DEVICE_FRAME_FN(lambda user, pl, st, w, hh, dev, fid, rc: fn([pl[i] for i in range(3)], [st[i] for i in range(3)],
                                                                          [w[i] for i in range(3)], [hh[i] for i in range(3)], dev, fid, rc)) if fn else DEVICE_FRAME_FN()
    lib().mp2v_decoder_set_device_renderer(h, self._dev_cb, None)


Decoder.set_device_renderer = _set_device_renderer


class ParsedPicture:
    def __init__(self, params, mb, coef, temporal_reference, gop):
        self.params, self.mb, self.coef, self.temporal_reference, self.gop = params, mb, coef, temporal_reference, gop


def parse_stream(padded, size, width, height, chroma_format, threads=1, want_records=True):
    """host-only slice parse -> (pictures, wall_seconds, cpu_seconds)"""
    L = lib()
    buf = np.ascontiguousarray(padded, np.uint8)
    h = C.c_void_p()
    err = C.create_string_buffer(512)
    rc = L.mp2v_parse_stream(buf.ctypes.data, size, width, height, chroma_format, threads, C.byref(h), err, 512)
    if rc != OK:
        raise ReconError("parse failed (%d): %s" % (rc, err.value.decode()))
    try:
        pics = []
        if want_records:
            n_mb = (width // 16) * (height // 16)
            for i in range(L.mp2v_parsed_num_pictures(h)):
                pp = PicParams()
                mb = C.POINTER(MbInfo)()
                coef = C.POINTER(C.c_uint32)()
                n_coef = C.c_uint32()
                tr = C.c_int32()
                gop = C.c_int32()
                L.mp2v_parsed_picture(h, i, C.byref(pp), C.byref(mb), C.byref(coef), C.byref(n_coef), C.byref(tr), C.byref(gop))
                mba = np.ctypeslib.as_array(C.cast(mb, C.POINTER(C.c_uint8)), shape=(n_mb * 16,)).copy().view(mb_dtype)
                ca = np.ctypeslib.as_array(coef, shape=(n_coef.value,)).copy() if n_coef.value else np.zeros(0, np.uint32)
                pics.append(ParsedPicture(pp, mba, ca, tr.value, gop.value))
        n = L.mp2v_parsed_num_pictures(h)
        return pics, L.mp2v_parsed_wall_seconds(h), L.mp2v_parsed_cpu_seconds(h), n
    finally:
        L.mp2v_parsed_free(h)
