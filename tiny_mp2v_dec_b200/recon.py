"""ctypes binding of the reconstruction C ABI (include/mp2v_recon.h, libmp2v_b200.so).

There is deliberately no fallback: if the library is missing it is built, if it cannot be loaded or
no sm_100a device is usable every call raises."""
import ctypes as C
import os

import numpy as np

from . import build as _build
from .abi import (OK, FrameLayout, MbInfo, PicParams, PicSyntax, Picture, ReconConfig, ReconStats, SliceRef, mb_dtype)

U8P = C.POINTER(C.c_uint8)
_lib = None

EXPORTS = [
    "mp2v_frame_layout", "mp2v_recon_create", "mp2v_recon_destroy", "mp2v_recon_last_error",
    "mp2v_recon_acquire_picture", "mp2v_recon_release_picture", "mp2v_recon_submit", "mp2v_recon_submit_slices", "mp2v_recon_stage_slices", "mp2v_recon_submit_staged", "mp2v_recon_stream_begin", "mp2v_recon_stream_codes", "mp2v_recon_stream_add", "mp2v_recon_submit_stream_picture", "mp2v_recon_precheck", "mp2v_recon_flush",
    "mp2v_recon_sync", "mp2v_recon_reset", "mp2v_recon_upload", "mp2v_recon_run_resident", "mp2v_recon_download_frame",
    "mp2v_recon_map_frame", "mp2v_recon_upload_frame", "mp2v_recon_frame_device_ptrs", "mp2v_recon_convert_frames", "mp2v_recon_convert_frame_nv12", "mp2v_recon_convert_frames_nv12", "mp2v_recon_wait_frame",
    "mp2v_recon_set_timing", "mp2v_recon_get_stats", "mp2v_recon_timer_start", "mp2v_recon_timer_stop",
    "mp2v_recon_numa_node", "mp2v_numa_parse_cpu_list",
]


class ByteRange(C.Structure):
    _fields_ = [("offset", C.c_size_t), ("bytes", C.c_size_t)]


class ReconError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        path = os.environ.get("MP2V_B200_LIB", _build.PRODUCT_LIB)     # dev knob: kernel-tuning variants (tools/dev)
        if not os.path.exists(path):
            _build.build_product()
        L = C.CDLL(path)
        P = C.POINTER
        L.mp2v_frame_layout.argtypes = [C.c_int, C.c_int, C.c_int, P(FrameLayout)]
        L.mp2v_recon_create.argtypes = [P(ReconConfig), P(C.c_void_p)]
        L.mp2v_recon_destroy.argtypes = [C.c_void_p]
        L.mp2v_recon_destroy.restype = None
        L.mp2v_recon_last_error.argtypes = [C.c_void_p]
        L.mp2v_recon_last_error.restype = C.c_char_p
        L.mp2v_recon_acquire_picture.argtypes = [C.c_void_p, P(P(Picture))]
        L.mp2v_recon_release_picture.argtypes = [C.c_void_p, P(Picture)]
        L.mp2v_recon_submit.argtypes = [C.c_void_p, P(Picture)]
        L.mp2v_recon_submit_slices.argtypes = [C.c_void_p, P(Picture), P(PicSyntax), P(SliceRef), C.c_int]
        L.mp2v_recon_stage_slices.argtypes = [C.c_void_p, P(Picture), P(PicSyntax), P(SliceRef), C.c_int]
        L.mp2v_recon_submit_staged.argtypes = [C.c_void_p, P(Picture)]
        L.mp2v_recon_stream_begin.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_int, P(P(C.c_uint32)), P(C.c_uint32)]
        L.mp2v_recon_submit_stream_picture.argtypes = [C.c_void_p, P(Picture), P(PicSyntax), P(C.c_uint32), C.c_int]
        L.mp2v_recon_precheck.argtypes = [C.c_void_p, P(Picture)]
        L.mp2v_recon_flush.argtypes = [C.c_void_p]
        L.mp2v_recon_sync.argtypes = [C.c_void_p]
        L.mp2v_recon_reset.argtypes = [C.c_void_p]
        L.mp2v_recon_upload.argtypes = [C.c_void_p, P(Picture)]
        L.mp2v_recon_run_resident.argtypes = [C.c_void_p, P(P(Picture)), P(C.c_int32), C.c_int]
        L.mp2v_recon_download_frame.argtypes = [C.c_void_p, C.c_int, U8P * 3, C.c_int32 * 3]
        L.mp2v_recon_map_frame.argtypes = [C.c_void_p, C.c_int, U8P * 3, C.c_int32 * 3]
        L.mp2v_recon_upload_frame.argtypes = [C.c_void_p, C.c_int, U8P * 3, C.c_int32 * 3]
        L.mp2v_recon_frame_device_ptrs.argtypes = [C.c_void_p, C.c_int, C.c_void_p * 3, C.c_int32 * 3]
        L.mp2v_recon_convert_frame_nv12.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int32]
        L.mp2v_recon_convert_frames_nv12.argtypes = [C.c_void_p, P(C.c_int32), P(C.c_void_p), C.c_int, C.c_int32]
        L.mp2v_recon_convert_frames.argtypes = [C.c_void_p, C.c_int, P(C.c_int32), P(C.c_void_p), C.c_int, C.c_int32]
        L.mp2v_recon_wait_frame.argtypes = [C.c_void_p, C.c_int]
        L.mp2v_recon_numa_node.argtypes = [C.c_void_p]
        L.mp2v_recon_set_timing.argtypes = [C.c_void_p, C.c_int]
        L.mp2v_recon_get_stats.argtypes = [C.c_void_p, P(ReconStats), C.c_int]
        L.mp2v_recon_timer_start.argtypes = [C.c_void_p]
        L.mp2v_recon_timer_stop.argtypes = [C.c_void_p, P(C.c_double)]
        _lib = L
    return _lib


def frame_layout(width, height, chroma_format):
    lay = FrameLayout()
    if lib().mp2v_frame_layout(width, height, chroma_format, C.byref(lay)) != OK:
        raise ReconError("bad geometry %dx%d cf=%d" % (width, height, chroma_format))
    return lay


class Recon:
    """One reconstruction context = one device."""

    def __init__(self, width, height, chroma_format, n_frames=8, n_pictures=8, device=0, max_batch=0, flags=1,
                 coef_capacity=0, bitstream_capacity=0):
        self.L = lib()
        cfg = ReconConfig(device, width, height, chroma_format, n_frames, n_pictures, max_batch, flags, coef_capacity,
                          bitstream_capacity)
        h = C.c_void_p()
        rc = self.L.mp2v_recon_create(C.byref(cfg), C.byref(h))
        if rc != OK:
            raise ReconError("mp2v_recon_create failed (%d): %s" % (rc, self.L.mp2v_recon_last_error(None).decode()))
        self.h = h
        self.width, self.height, self.chroma_format = width, height, chroma_format
        self.lay = frame_layout(width, height, chroma_format)

    def close(self):
        if self.h:
            self.L.mp2v_recon_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != OK:
            raise ReconError("mp2v_recon error %d: %s" % (rc, self.L.mp2v_recon_last_error(self.h).decode()))

    # ---- pictures
    def acquire(self):
        p = C.POINTER(Picture)()
        self._ck(self.L.mp2v_recon_acquire_picture(self.h, C.byref(p)))
        return p

    def fill(self, pic, params, mb, coef, dst, l0=-1, l1=-1):
        """copy ground-truth / parsed records into an acquired picture's pinned buffers"""
        p = pic.contents
        n_mb, n_coef = len(mb), len(coef)
        assert n_mb == p.mb_count, (n_mb, p.mb_count)
        if n_coef > p.coef_capacity:
            raise ReconError("picture needs %d coefficient records, slot holds %d" % (n_coef, p.coef_capacity))
        C.memmove(p.params, C.byref(params), C.sizeof(PicParams))
        p.params.contents.dst_frame, p.params.contents.l0_frame, p.params.contents.l1_frame = dst, l0, l1
        p.params.contents.n_coef = n_coef
        C.memmove(p.mb, mb.ctypes.data, n_mb * 16)
        if n_coef:
            C.memmove(p.coef, np.ascontiguousarray(coef, np.uint32).ctypes.data, n_coef * 4)

    def submit(self, pic):
        self._ck(self.L.mp2v_recon_submit(self.h, pic))

    def submit_slices(self, pic, params, data, slices, f_code, intra_dc_precision=0, q_scale_type=0, intra_vlc_format=1,
                      dst=0, l0=-1, l1=-1, field_dct_syntax=0):
        """device-side parsing (contexts created with flags | RECON_DEVICE_VLC): data = uint8 array holding the
        coded picture, slices = [(payload offset in data, payload bytes, slice_start_code value), ...]"""
        p = pic.contents
        C.memmove(p.params, C.byref(params), C.sizeof(PicParams))
        p.params.contents.dst_frame, p.params.contents.l0_frame, p.params.contents.l1_frame = dst, l0, l1
        sy = PicSyntax(((C.c_int32 * 2) * 2)((C.c_int32 * 2)(*f_code[0]), (C.c_int32 * 2)(*f_code[1])),
                       intra_dc_precision, q_scale_type, intra_vlc_format, field_dct_syntax)
        buf = np.ascontiguousarray(data, np.uint8)
        refs = (SliceRef * max(len(slices), 1))(*[SliceRef(buf.ctypes.data + off, n, code) for off, n, code in slices])
        self._ck(self.L.mp2v_recon_submit_slices(self.h, pic, C.byref(sy), refs, len(slices)))

    def stream_begin(self, data, size, scan=True, part=None, deferred=False):
        """stream-resident front end: copy data[:size] (uint8 array, kept alive by the caller until the pictures are
        submitted) to the device; with scan -> ascending offsets of every 00 00 01 prefix (numpy copy).
        part = (offset, bytes): copy and scan only that part (offset a multiple of 16); deferred: launch only, the
        list comes from stream_codes()"""
        self._stream = np.ascontiguousarray(data, np.uint8)
        codes = C.POINTER(C.c_uint32)()
        n = C.c_uint32()
        rng = ByteRange(part[0], part[1]) if part is not None else None
        mode = 0 if not scan else 2 if deferred else 1
        self._ck(self.L.mp2v_recon_stream_begin(self.h, self._stream.ctypes.data, size, C.byref(rng) if rng is not None else None,
                                                1 if rng is not None else 0, mode, C.byref(codes) if mode == 1 else None,
                                                C.byref(n) if mode == 1 else None))
        if mode != 1:
            return None
        return np.ctypeslib.as_array(codes, shape=(n.value,)).copy() if n.value else np.zeros(0, np.uint32)

    def stream_codes(self):
        """the list of a scan launched with stream_begin(..., deferred=True)"""
        codes = C.POINTER(C.c_uint32)()
        n = C.c_uint32()
        self._ck(self.L.mp2v_recon_stream_codes(self.h, C.byref(codes), C.byref(n)))
        return np.ctypeslib.as_array(codes, shape=(n.value,)).copy() if n.value else np.zeros(0, np.uint32)

    def stream_add(self, ranges):
        """copy further (offset, bytes) ranges of the resident stream"""
        arr = (ByteRange * max(len(ranges), 1))(*[ByteRange(o, b) for o, b in ranges])
        self._ck(self.L.mp2v_recon_stream_add(self.h, arr, len(ranges)))

    def submit_stream_picture(self, pic, params, slice_offsets, f_code, intra_dc_precision=0, q_scale_type=0, intra_vlc_format=1,
                              dst=0, l0=-1, l1=-1, field_dct_syntax=0):
        """slice_offsets: byte offsets of the slices' START CODES in the resident stream"""
        p = pic.contents
        C.memmove(p.params, C.byref(params), C.sizeof(PicParams))
        p.params.contents.dst_frame, p.params.contents.l0_frame, p.params.contents.l1_frame = dst, l0, l1
        sy = PicSyntax(((C.c_int32 * 2) * 2)((C.c_int32 * 2)(*f_code[0]), (C.c_int32 * 2)(*f_code[1])),
                       intra_dc_precision, q_scale_type, intra_vlc_format, field_dct_syntax)
        offs = (C.c_uint32 * max(len(slice_offsets), 1))(*[int(o) for o in slice_offsets])
        self._ck(self.L.mp2v_recon_submit_stream_picture(self.h, pic, C.byref(sy), offs, len(slice_offsets)))

    def upload(self, pic):
        self._ck(self.L.mp2v_recon_upload(self.h, pic))

    def release(self, pic):
        self._ck(self.L.mp2v_recon_release_picture(self.h, pic))

    def run_resident(self, pics, levels=None):
        n = len(pics)
        arr = (C.POINTER(Picture) * n)(*pics)
        lv = (C.c_int32 * n)(*levels) if levels is not None else None
        self._ck(self.L.mp2v_recon_run_resident(self.h, arr, lv, n))

    def flush(self):
        self._ck(self.L.mp2v_recon_flush(self.h))

    def sync(self):
        self._ck(self.L.mp2v_recon_sync(self.h))

    def reset(self):
        self._ck(self.L.mp2v_recon_reset(self.h))

    # ---- frames
    def download(self, frame_id):
        """-> cropped planar YUV bytes of a frame"""
        planes = [np.empty((self.lay.height[p], self.lay.width[p]), np.uint8) for p in range(3)]
        ptrs = (U8P * 3)(*[a.ctypes.data_as(U8P) for a in planes])
        strides = (C.c_int32 * 3)(*[self.lay.width[p] for p in range(3)])
        self._ck(self.L.mp2v_recon_download_frame(self.h, frame_id, ptrs, strides))
        return b"".join(a.tobytes() for a in planes)

    def upload_frame(self, frame_id, planes):
        planes = [np.ascontiguousarray(a, np.uint8) for a in planes]
        ptrs = (U8P * 3)(*[a.ctypes.data_as(U8P) for a in planes])
        strides = (C.c_int32 * 3)(*[a.shape[1] for a in planes])
        self._ck(self.L.mp2v_recon_upload_frame(self.h, frame_id, ptrs, strides))

    def convert_nv12(self, frame_id, dst_device_ptr, dst_pitch):
        """frame -> NV12 in a device buffer of the caller (e.g. a torch CUDA tensor's data_ptr()); asynchronous"""
        self._ck(self.L.mp2v_recon_convert_frame_nv12(self.h, frame_id, C.c_void_p(dst_device_ptr), dst_pitch))

    def convert_nv12_batch(self, frame_ids, dst_device_ptrs, dst_pitch):
        """several frames, 32 per launch"""
        n = len(frame_ids)
        ids = (C.c_int32 * n)(*frame_ids)
        ptrs = (C.c_void_p * n)(*dst_device_ptrs)
        self._ck(self.L.mp2v_recon_convert_frames_nv12(self.h, ids, ptrs, n, dst_pitch))

    OUT_FORMATS = {"nv12": 0, "p010": 1, "uyvy": 2}

    def convert_batch(self, fmt, frame_ids, dst_device_ptrs, dst_pitch):
        """frames -> NV12 / P010 (4:2:0) or UYVY (4:2:2) in device buffers of the caller; asynchronous, 32 frames per launch"""
        n = len(frame_ids)
        ids = (C.c_int32 * n)(*frame_ids)
        ptrs = (C.c_void_p * n)(*dst_device_ptrs)
        self._ck(self.L.mp2v_recon_convert_frames(self.h, self.OUT_FORMATS[fmt], ids, ptrs, n, dst_pitch))

    def wait_frame(self, frame_id):
        self._ck(self.L.mp2v_recon_wait_frame(self.h, frame_id))

    # ---- statistics
    def numa_node(self):
        """NUMA node the context's pinned memory was placed on, -1 on single-node hosts"""
        return int(self.L.mp2v_recon_numa_node(self.h))

    def set_timing(self, on=True):
        self._ck(self.L.mp2v_recon_set_timing(self.h, 1 if on else 0))

    def timer_start(self):
        self._ck(self.L.mp2v_recon_timer_start(self.h))

    def timer_stop(self):
        """-> device milliseconds since timer_start() on the compute stream (waits for the work)"""
        ms = C.c_double()
        self._ck(self.L.mp2v_recon_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def stats(self, reset=False):
        s = ReconStats()
        self._ck(self.L.mp2v_recon_get_stats(self.h, C.byref(s), 1 if reset else 0))
        return s


def reconstruct_stream(stream, recon=None, n_frames=None, batch=True):
    """Reconstruct every picture of a generated Stream from its ground-truth records on the GPU.
    Frame id = coded index (the pool is sized to hold the whole stream).  Returns cropped planar YUV in
    display order -- the CUDA counterpart of tests/oracle_lib.oracle_decode_stream."""
    n = len(stream.pictures)
    own = recon is None
    if own:
        recon = Recon(stream.width, stream.height, stream.chroma_format, n_frames=n_frames or n, n_pictures=min(n, 8))
    try:
        for idx, pic in enumerate(stream.pictures):
            h = recon.acquire()
            recon.fill(h, pic.params, pic.mb, pic.coef, dst=idx, l0=pic.params.l0_frame, l1=pic.params.l1_frame)
            recon.submit(h)
            if not batch:
                recon.flush()
        recon.sync()
        return b"".join(recon.download(i) for i in stream.display_order())
    finally:
        if own:
            recon.close()
